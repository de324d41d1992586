"""iron_b200: the surface-rendering hot path of IRON (sphere tracing + bisection against the neural SDF,
IDR implicit differentiation, colocated-flash GGX shading, forward and backward) as hand-written sm_100a
CUDA kernels behind the reference's Python module API.  CUDA only; there is no CPU fallback."""
from . import _lib
from .fields import RenderingNetwork, SDFNetwork
from .network_conf import (PointLightNetwork, choose_renderer, init_rendering_network_dict, init_sdf_network_dict)
from .raytracer import (Camera, RayTracer, intersect_sphere, locate_edge_points, raytrace_camera, raytrace_pixels,
                        render_camera, render_edge_pixels, render_normal_and_color, reparam_points)
from .render_fn import make_render_fn, make_render_fn_comp2
from .renderer_ggx import CompositeRenderer, GGXColocatedRenderer
from .rendering_func import get_materials, get_materials_comp
from .embedder import get_embedder
from .optim import FusedAdam
from .image_losses import PyramidL2Loss, ssim_loss_fn
from .renderer import NeRF, NeuSRenderer, SingleVarianceNetwork, sample_pdf
from .step import GraphedNeusStep, GraphedStage2Step, stage2_step

__all__ = [
    "SDFNetwork", "RenderingNetwork", "PointLightNetwork", "GGXColocatedRenderer", "RayTracer", "Camera",
    "intersect_sphere", "raytrace_pixels", "raytrace_camera", "render_camera", "render_normal_and_color",
    "reparam_points", "locate_edge_points", "render_edge_pixels", "get_materials", "get_embedder", "make_render_fn", "stage2_step", "GraphedStage2Step", "FusedAdam",
    "PyramidL2Loss", "ssim_loss_fn", "CompositeRenderer", "get_materials_comp", "make_render_fn_comp2",
    "init_sdf_network_dict", "init_rendering_network_dict", "choose_renderer",
    "NeuSRenderer", "GraphedNeusStep", "NeRF", "SingleVarianceNetwork", "sample_pdf",
]
