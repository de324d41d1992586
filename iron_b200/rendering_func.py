"""get_materials with the reference's signature (models/rendering_func.py:5-16): the three material
RenderingNetworks (fused CUDA forward/backward) plus the reference's post-ops (abs, channel mean, +0.01)."""
import torch


def get_materials(network_dict, points, normals, features, is_metal=False):
    diffuse_albedo = network_dict["diffuse_albedo_network"](points, normals, -normals, features).abs()
    specular_albedo = network_dict["specular_albedo_network"](points, normals, None, features).abs()
    if not is_metal:
        specular_albedo = torch.mean(specular_albedo, dim=-1, keepdim=True).expand_as(specular_albedo)
    specular_roughness = network_dict["specular_roughness_network"](points, normals, None, features).abs() + 0.01
    return {
        "diffuse_albedo": diffuse_albedo,
        "specular_albedo": specular_albedo,
        "specular_roughness": specular_roughness,
    }
