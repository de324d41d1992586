"""The 'ggx' render_fn of the stage-2 driver (render_surface.py:117-156): normalise the SDF gradient, query the
three material networks, shade with the colocated-flash GGX model and scatter the hit results into dense,
zero-filled buffers.  `make_render_fn(renderer)` returns the callback `render_normal_and_color` expects."""
import torch

from . import _fused
from .rendering_func import get_materials, get_materials_comp


def _scatter(dense_shape, idx, src, width):
    if idx is None:            # dense shading: one result per ray, already in ray order
        return src.reshape(dense_shape) if width == 1 else src.reshape(list(dense_shape) + [width])
    n = 1
    for s in dense_shape:
        n *= s
    if width == 1:
        out = torch.zeros(n, dtype=torch.float32, device=src.device).index_copy(0, idx, src.reshape(-1))
        return out.reshape(dense_shape)
    out = torch.zeros(n, width, dtype=torch.float32, device=src.device).index_copy(0, idx, src)
    return out.reshape(list(dense_shape) + [width])


def make_render_fn(renderer, is_metal=False):
    def render_fn(interior_mask, color_network_dict, ray_o, ray_d, points, normals, features):
        dots_sh = list(interior_mask.shape)
        dev = interior_mask.device
        dense = bool(getattr(interior_mask, "_ironb_dense", False))    # every ray is shaded: the scatter is the identity
        idx = getattr(interior_mask, "_ironb_idx", None)
        if idx is None and not dense:
            idx = torch.nonzero(interior_mask.reshape(-1), as_tuple=False).reshape(-1)
        z3 = lambda: torch.zeros(dots_sh + [3], dtype=torch.float32, device=dev)
        if points.shape[0] == 0:
            return {"color": z3(), "diffuse_color": z3(), "specular_color": z3(), "diffuse_albedo": z3(),
                    "specular_albedo": z3(), "specular_roughness": z3()[..., 0].clone(), "normal": z3()}
        # n = g / (|g| + 1e-10) and |x - o| in one launch (render_surface.py:135-146)
        normals, dist = _fused.unit_normal_and_distance(normals, points, ray_o)
        params = get_materials(network_dict=color_network_dict, points=points, normals=normals, features=features,
                               is_metal=is_metal)
        res = renderer(color_network_dict["point_light_network"](), dist, normals, -ray_d, params=params)
        if dense:
            idx = None
        return {
            "color": _scatter(dots_sh, idx, res["rgb"], 3),
            "diffuse_color": _scatter(dots_sh, idx, res["diffuse_rgb"], 3),
            "specular_color": _scatter(dots_sh, idx, res["specular_rgb"], 3),
            "diffuse_albedo": _scatter(dots_sh, idx, params["diffuse_albedo"], 3),
            "specular_albedo": _scatter(dots_sh, idx, params["specular_albedo"].contiguous(), 3),
            "specular_roughness": _scatter(dots_sh, idx, params["specular_roughness"], 1),
            "normal": _scatter(dots_sh, idx, normals, 3),
        }

    return render_fn


def make_render_fn_comp2(renderer):
    """The 'comp2' render_fn of the fork's working driver (model_bed.py:227-298, with its get_material_comp :115-154): the eight
    material heads + the env-light head, CompositeRenderer shading, scatter into dense zero-filled buffers.  Shapes as there:
    specular_roughness / metallic_eta / metallic_k / dielectric_eta / metallic / dielectric / costheta are [..., 1], env_light
    is the 1-channel head broadcast to [..., 3]."""
    def render_fn(interior_mask, color_network_dict, ray_o, ray_d, points, normals, features):
        dots_sh = list(interior_mask.shape)
        dev = interior_mask.device
        dense = bool(getattr(interior_mask, "_ironb_dense", False))
        idx = getattr(interior_mask, "_ironb_idx", None)
        if idx is None and not dense:
            idx = torch.nonzero(interior_mask.reshape(-1), as_tuple=False).reshape(-1)
        z = lambda w: torch.zeros(dots_sh + [w], dtype=torch.float32, device=dev)
        keys3 = ("color", "diffuse_color", "specular_color", "diffuse_albedo", "specular_albedo", "normal", "metallic_rgb",
                 "dielectric_rgb", "env_light")
        keys1 = ("specular_roughness", "metallic_eta", "metallic_k", "dielectric_eta", "metallic", "dielectric", "costheta")
        if points.shape[0] == 0:
            out = {k: z(3) for k in keys3}
            out.update({k: z(1) for k in keys1})
            return out
        normals = normals / (normals.norm(dim=-1, keepdim=True) + 1e-10)
        ray_d_norm = ray_d / (ray_d.norm(dim=-1, keepdim=True) + 1e-10)
        costheta = torch.sum(-1 * ray_d_norm * normals, dim=-1, keepdim=True)
        params = get_materials_comp(color_network_dict, points, normals, features)
        if "env_light_network" in color_network_dict:
            params["env_light"] = color_network_dict["env_light_network"](points, None, None, features).abs()
        res = renderer(color_network_dict["point_light_network"](), (points - ray_o).norm(dim=-1, keepdim=True), normals, -ray_d,
                       params=params)
        if dense:
            idx = None
        sc3 = lambda t: _scatter(dots_sh, idx, t.contiguous(), 3)
        sc1 = lambda t: _scatter(dots_sh, idx, t.contiguous(), 1).reshape(dots_sh + [1])
        out = {"color": sc3(res["rgb"]), "diffuse_color": sc3(res["diffuse_rgb"]), "specular_color": sc3(res["specular_rgb"]),
               "diffuse_albedo": sc3(params["diffuse_albedo"]), "specular_albedo": sc3(params["specular_albedo"]),
               "normal": sc3(normals), "metallic_rgb": sc3(res["metallic_rgb"]), "dielectric_rgb": sc3(res["dielectric_rgb"]),
               "specular_roughness": sc1(params["specular_roughness"]), "metallic_eta": sc1(params["metallic_eta"]),
               "metallic_k": sc1(params["metallic_k"]), "dielectric_eta": sc1(params["dielectric_eta"]),
               "metallic": sc1(params["metallic"]), "dielectric": sc1(params["dielectric"]), "costheta": sc1(costheta)}
        env = params.get("env_light")
        out["env_light"] = sc3(env.expand(-1, 3)) if env is not None else z(3)
        return out

    return render_fn
