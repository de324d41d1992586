"""get_materials with the reference's signature (models/rendering_func.py:5-16): the three material
RenderingNetworks (fused CUDA forward/backward) plus the reference's post-ops (abs, channel mean, +0.01)."""
import torch


def get_materials(network_dict, points, normals, features, is_metal=False):
    """network_dict may carry `"_ironb_streams": [s1, s2]` (GraphedStage2Step sets it): the specular-albedo and roughness
    networks then run on those streams next to the diffuse one.  The three MLPs are independent and each of their GEMMs
    fills only 64 of the 148 SMs (256 output features), so two run side by side; autograd replays the same placement in
    the backward pass.  Same arithmetic either way."""
    streams = network_dict.get("_ironb_streams") if isinstance(network_dict, dict) else None

    def diffuse():
        return network_dict["diffuse_albedo_network"](points, normals, -normals, features).abs()

    def specular():
        sa = network_dict["specular_albedo_network"](points, normals, None, features).abs()
        if not is_metal:
            sa = torch.mean(sa, dim=-1, keepdim=True).expand_as(sa)
        return sa

    def roughness():
        return network_dict["specular_roughness_network"](points, normals, None, features).abs() + 0.01

    if streams and points.is_cuda:
        cur = torch.cuda.current_stream(points.device)
        for net in ("diffuse_albedo_network", "specular_albedo_network", "specular_roughness_network"):
            network_dict[net].folded()                    # folds happen on the current stream, before the fork
        for st in streams[:2]:
            st.wait_stream(cur)
        with torch.cuda.stream(streams[0]):
            specular_albedo = specular()
        with torch.cuda.stream(streams[1]):
            specular_roughness = roughness()
        diffuse_albedo = diffuse()
        for st in streams[:2]:
            cur.wait_stream(st)
    else:
        diffuse_albedo = diffuse()
        specular_albedo = specular()
        specular_roughness = roughness()
    return {
        "diffuse_albedo": diffuse_albedo,
        "specular_albedo": specular_albedo,
        "specular_roughness": specular_roughness,
    }


def get_materials_comp(network_dict, points, normals, features):
    """models/rendering_func.py:19-48: the nine material heads of the "comp2" renderer (all RenderingNetworks, fused CUDA
    forward / backward), each followed by the reference's abs()."""
    net = lambda name, v: network_dict[name](points, normals, v, features).abs()
    return {
        "diffuse_albedo": net("diffuse_albedo_network", -normals),
        "specular_albedo": net("specular_albedo_network", None),
        "metallic": net("metallic_network", None),
        "dielectric": net("dielectric_network", None),
        "specular_roughness": net("specular_roughness_network", None),
        "metallic_eta": net("metallic_eta_network", None),
        "metallic_k": net("metallic_k_network", None),
        "dielectric_eta": net("dielectric_eta_network", None),
    }
