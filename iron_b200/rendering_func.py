"""get_materials with the reference's signature (models/rendering_func.py:5-16): the three material
RenderingNetworks (fused CUDA forward/backward) plus the reference's post-ops (abs, channel mean, +0.01)."""
import torch

from . import _fused


def get_materials(network_dict, points, normals, features, is_metal=False):
    """network_dict may carry `"_ironb_streams": [s1, s2]` (GraphedStage2Step sets it): the specular-albedo and roughness
    networks then run on those streams next to the diffuse one.  The three MLPs are independent and each of their GEMMs
    fills only 64 of the 148 SMs (256 output features), so two run side by side; autograd replays the same placement in
    the backward pass.  Same arithmetic either way."""
    streams = network_dict.get("_ironb_streams") if isinstance(network_dict, dict) else None
    d_net, s_net, r_net = (network_dict[k] for k in ("diffuse_albedo_network", "specular_albedo_network",
                                                      "specular_roughness_network"))
    neg_n = -normals
    if streams and points.is_cuda:
        cur = torch.cuda.current_stream(points.device)
        for net in (d_net, s_net, r_net):
            net.folded()                                  # folds happen on the current stream, before the fork
        for st in streams[:2]:
            st.wait_stream(cur)
        with torch.cuda.stream(streams[0]):
            sa = s_net(points, normals, None, features)
        with torch.cuda.stream(streams[1]):
            sr = r_net(points, normals, None, features)
        da = d_net(points, normals, neg_n, features)
        for st in streams[:2]:
            cur.wait_stream(st)
    else:
        da = d_net(points, normals, neg_n, features)
        sa = s_net(points, normals, None, features)
        sr = r_net(points, normals, None, features)
    if points.is_cuda and da.dim() == 2 and da.shape[-1] == 3 and sa.shape[-1] == 3 and sr.shape[-1] == 1:
        # abs / channel mean / + 0.01 of all three heads in one launch (csrc/glue.cu); same values as the expressions below
        diffuse_albedo, specular_albedo, specular_roughness = _fused.material_post(da, sa, sr, is_metal)
    else:
        diffuse_albedo = da.abs()
        specular_albedo = sa.abs()
        if not is_metal:
            specular_albedo = torch.mean(specular_albedo, dim=-1, keepdim=True).expand_as(specular_albedo)
        specular_roughness = sr.abs() + 0.01
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
