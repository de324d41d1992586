"""Golden vectors for CompositeRenderer.forward from the REAL reference class (models/renderer_ggx.py:520-858),
run on CPU in the build container.  TEST INFRASTRUCTURE ONLY.        python oracle/make_golden_composite.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG   # noqa: E402


def main():
    fields, raytracer, renderer_ggx, rendering_func, network_conf = MG.import_reference()
    cwd = os.getcwd()
    os.chdir("/root/reference")          # the ctor globs ./resource/ior/*.spd (absent from the tree: empty lists either way)
    try:
        rend = renderer_ggx.CompositeRenderer(use_cuda=False)
    finally:
        os.chdir(cwd)
    g = torch.Generator().manual_seed(21)
    M = 96
    cosv = torch.rand(M, 1, generator=g) * 0.98 + 0.01
    v = torch.nn.functional.normalize(torch.randn(M, 3, generator=g), dim=-1)
    t = torch.nn.functional.normalize(torch.cross(v, torch.randn(M, 3, generator=g), dim=-1), dim=-1)
    n = cosv * v + torch.sqrt(1 - cosv * cosv) * t
    n[0] = v[0]                       # head-on: cos clamps at 0.99999
    n[1] = -v[1]                      # back-facing: cos clamps at 1e-5
    n[2] = t[2]                       # grazing
    P = {
        "diffuse_albedo": torch.rand(M, 3, generator=g),
        "specular_albedo": torch.rand(M, 3, generator=g),
        "specular_roughness": torch.rand(M, 1, generator=g) * 0.99 + 0.01,
        "metallic": torch.rand(M, 1, generator=g),
        "dielectric": torch.rand(M, 1, generator=g),
        "metallic_eta": torch.rand(M, 1, generator=g) * 5.45 + 0.05,      # crosses both clamps (0.099999, 4.999999)
        "metallic_k": torch.rand(M, 1, generator=g) * 10.95 + 0.05,       # (0.099999, 9.999999)
        "dielectric_eta": torch.rand(M, 1, generator=g) * 1.2 + 0.9,      # (1.000001, 1.999999)
    }
    P["diffuse_albedo"][3] = 0.0      # below the 1e-5 clamp
    P["specular_albedo"][4] = 0.0
    P["specular_roughness"][5] = 0.0
    dist = torch.rand(M, 1, generator=g) + 1.5
    light = torch.tensor(32.0)
    leaves = {k: x.clone().requires_grad_(True) for k, x in P.items()}
    light_l, dist_l, n_l = light.clone().requires_grad_(True), dist.clone().requires_grad_(True), n.clone().requires_grad_(True)
    out = rend(light_l, dist_l, n_l, v, params=leaves)
    ups = {k: torch.randn(M, 3, generator=g) for k in ("rgb", "specular_rgb", "metallic_rgb", "dielectric_rgb", "diffuse_rgb")}
    loss = sum((out[k] * ups[k]).sum() for k in ups)
    names = ["diffuse_albedo", "specular_albedo", "specular_roughness", "metallic_eta", "metallic_k", "dielectric_eta"]
    grads = torch.autograd.grad(loss, [light_l, dist_l, n_l] + [leaves[k] for k in names], allow_unused=True)
    data = {"light": light.numpy(), "dist": dist.numpy(), "normal": n.numpy(), "viewdir": v.numpy()}
    data.update({"p." + k: x.numpy() for k, x in P.items()})
    data.update({"out." + k: out[k].detach().numpy() for k in out})
    data.update({"up." + k: x.numpy() for k, x in ups.items()})
    for k, gr in zip(["light", "dist", "normal"] + names, grads):
        data["g." + k] = gr.numpy()
    data["diffuse_is_rgb"] = np.array(out["diffuse_rgb"].data_ptr() == out["rgb"].data_ptr())
    path = os.path.join(HERE, "..", "tests", "golden", "composite.npz")
    np.savez_compressed(path, **data)
    print("wrote", path, os.path.getsize(path), "bytes; diffuse_rgb is rgb:", bool(data["diffuse_is_rgb"]))


if __name__ == "__main__":
    main()
