"""One full stage-2 step (trace -> get_all -> reparam -> materials -> GGX -> loss -> backward) on the CUDA path vs
the golden step produced by the reference modules (32x32 silhouette crop, H=256)."""
import numpy as np
import pytest
import torch

from oracle import iron_oracle as O
from util import T, TOL_GRAD_REL, TOL_NORMAL, TOL_RGB, assert_close, perturb, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

# The reference's own reproducibility floor (oracle/measure_floor.py -> tests/golden/floor.json): the fraction of hit normals
# on which the reference algorithm in fp32 agrees with itself in fp64 within 1e-4.  Two fp32-grade evaluations (the CUDA path
# and the reference's fp32) each deviate from the exact result, so the CUDA path is held to 2.5x the floor's outlier rate --
# for EVERY tracer mode, no per-mode slack.
import json as _json
import os as _os
FLOOR = _json.load(open(_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden", "floor.json")))


def normal_need(key):
    return 1.0 - 2.5 * (1.0 - FLOOR[key]["normals_within_1e-4"])


def build():
    import iron_b200
    torch.manual_seed(0)
    nets = iron_b200.init_rendering_network_dict("ggx")       # same RNG order as make_golden section 4
    torch.manual_seed(0)
    sdf = iron_b200.SDFNetwork(d_in=3, d_out=257, d_hidden=256, n_layers=8, skip_in=[4], multires=6, bias=0.5,
                               scale=1.0, geometric_init=True, weight_norm=True)
    perturb(sdf, 0.005, seed=1)
    sdf = sdf.to(DEV)
    nets["point_light_network"].set_light(32.0)
    K = torch.tensor(O.FIXTURE_K, dtype=torch.float64).reshape(4, 4).float().to(DEV)
    W2C = torch.tensor(O.FIXTURE_W2C, dtype=torch.float64).reshape(4, 4).float().to(DEV)
    cam512 = iron_b200.Camera(512, 512, K, W2C)
    return iron_b200, sdf, nets, cam512


def test_step_golden(golden, gemm_mode, trace_mode):
    g = golden("step_h256")
    ib, sdf, nets, cam512 = build()
    # the material nets were built after a color_network + under the same seed as the reference: check one pin
    cam, _, _ = cam512.crop_region(32, 32, ul_corner=tuple(int(v) for v in g["ul"]))
    rend = ib.GGXColocatedRenderer(use_cuda=True)
    loss, res = ib.stage2_step(sdf, nets, ib.RayTracer(), ib.make_render_fn(rend), cam, T(g["target"]).to(DEV),
                               T(g["eik_points"]).to(DEV), eik_weight=0.1)
    m = res["convergent_mask"].cpu().numpy()
    mr = g["mask"]
    assert (m == mr).mean() >= 0.998, f"hit mask agreement {(m == mr).mean():.4f} on {m.size} rays"
    both = m & mr
    # BASELINE tolerances per output, on common hits: normal 1e-4, RGB 1e-3, depth 1e-4 (quantile, see test_trace_gpu)
    # normals are evaluated where each tracer converged inside the +-5e-5 SDF band; on this silhouette crop the
    # converged points differ by up to ~1e-4 along grazing rays, which moves the normal by curvature x offset:
    # stated as >= 99 % of common hits within 1e-4 and all within 2e-3 (the same-point normal parity, <= 5e-5,
    # is asserted in test_sdf_gpu.py)
    nerr = np.abs(res["normal"].detach().cpu().numpy()[both] - g["res.normal"][both]).max(axis=-1)
    derr = np.abs(res["distance"].cpu().numpy()[both] - g["res.distance"][both])
    print(f"[{gemm_mode}/{trace_mode}] hits {int(m.sum())}/{int(mr.sum())} mask-agree {(m == mr).mean():.4f}  normals within 1e-4: "
          f"{100 * (nerr <= 1e-4).mean():.2f}% (max {nerr.max():.2e})  distance within 1e-4: {100 * (derr <= 1e-4).mean():.2f}% "
          f"(max {derr.max():.2e})")
    assert_close(res["normal"].detach().cpu().numpy()[both], g["res.normal"][both], TOL_NORMAL, what="normal",
                 frac=normal_need("h256_golden_crop"))
    assert_close(res["normal"].detach().cpu().numpy()[both], g["res.normal"][both], 2e-3, what="normal (all)")
    for k in ("color", "diffuse_color", "specular_color"):
        assert_close(res[k].detach().cpu().numpy()[both], g["res." + k][both], TOL_RGB, 1e-3, what=k, frac=0.995)
    for k in ("diffuse_albedo", "specular_albedo", "specular_roughness"):
        assert_close(res[k].detach().cpu().numpy()[both], g["res." + k][both], 1e-4, 1e-4, what=k, frac=0.995)
    assert_close(res["distance"].cpu().numpy()[both], g["res.distance"][both], 1e-4, what="distance", frac=0.995)
    same_mask = bool((m == mr).all())
    assert abs(float(loss) - float(g["loss"])) <= (1e-3 if same_mask else 2e-2) * abs(float(g["loss"])), (float(loss), float(g["loss"]))
    # parameter gradients: 1e-3 relative (L2 norm of every tensor) when the hit sets are identical; a ray that
    # flips adds/removes a whole pixel's contribution, so the bound is loosened to 5 % in that case and reported
    tol = TOL_GRAD_REL if same_mask else 5e-2
    named = [("sdf." + k, p) for k, p in sdf.named_parameters()]
    for nm in ("diffuse_albedo_network", "specular_albedo_network", "specular_roughness_network", "point_light_network"):
        named += [(nm + "." + k, p) for k, p in nets[nm].named_parameters()]
    worst = 0.0
    for k, p in named:
        assert p.grad is not None, k
        gr = p.grad.double().cpu()
        ref = g["gsum." + k]
        nrm_err = abs(gr.pow(2).sum().sqrt().item() - ref[2]) / max(ref[2], 1e-12)
        worst = max(worst, nrm_err)
        assert nrm_err <= tol, (k, gr.pow(2).sum().sqrt().item(), ref[2], "same_mask", same_mask)
        if "g." + k in g:
            r = rel_l2(gr.numpy(), g["g." + k])
            worst = max(worst, r)
            assert r <= tol * 2, (k, r, "same_mask", same_mask)
    print(f"step parity: same_mask={same_mask} worst relative gradient error {worst:.2e}")


def test_render_camera_inference(golden, gemm_mode):
    """is_training=False path: detached outputs, no autograd graph, same images."""
    g = golden("step_h256")
    ib, sdf, nets, cam512 = build()
    cam, _, _ = cam512.crop_region(32, 32, ul_corner=tuple(int(v) for v in g["ul"]))
    rend = ib.GGXColocatedRenderer(use_cuda=True)
    res = ib.render_camera(cam, sdf, ib.RayTracer(), nets, ib.make_render_fn(rend), fill_holes=False, handle_edges=False,
                           is_training=False)
    assert not res["color"].requires_grad
    m, mr = res["convergent_mask"].cpu().numpy(), g["mask"]
    both = m & mr
    assert_close(res["color"].cpu().numpy()[both], g["res.color"][both], TOL_RGB, 1e-3, what="color", frac=0.995)
    assert res["color"].shape == (32, 32, 3) and res["specular_roughness"].shape == (32, 32)
    assert float(res["color"][~res["convergent_mask"]].abs().max()) == 0.0


def test_step_with_edges_golden(golden, trace_mode):
    """The reference drivers' default step: render_camera(fill_holes=True, handle_edges=True, is_training=True) + loss +
    backward, against the golden step produced by the real reference (kornia's closing / sobel shared as restatements)."""
    g = golden("step_edges_h256")
    ib, sdf, nets, cam512 = build()
    cam, _, _ = cam512.crop_region(32, 32, ul_corner=tuple(int(v) for v in g["ul"]))
    rend = ib.GGXColocatedRenderer(use_cuda=True)
    loss, res = ib.stage2_step(sdf, nets, ib.RayTracer(), ib.make_render_fn(rend), cam, T(g["target"]).to(DEV),
                               T(g["eik_points"]).to(DEV), eik_weight=0.1, fill_holes=True, handle_edges=True)
    m, em = res["convergent_mask"].cpu().numpy(), res["edge_mask"].cpu().numpy()
    mr, emr = g["mask"], g["edge_mask"]
    print(f"[{trace_mode}] hits {int(m.sum())}/{int(mr.sum())}  edge pixels {int(em.sum())}/{int(emr.sum())}  "
          f"mask agree {(m == mr).mean():.4f}  edge-mask agree {(em == emr).mean():.4f}  loss {float(loss):.6f}/{float(g['loss']):.6f}")
    assert (m == mr).mean() >= 0.995 and (em == emr).mean() >= 0.995
    same = bool((m == mr).all() and (em == emr).all())
    got_pix = set(res["edge_pixel_idx"].cpu().numpy().tolist())
    ref_pix = set(g["edge_pixel_idx"].tolist())
    assert len(got_pix & ref_pix) >= 0.9 * len(ref_pix)
    both = (m | em) & (mr | emr) & (em == emr)
    assert_close(res["color"].detach().cpu().numpy()[both], g["res.color"][both], TOL_RGB, 1e-3, what="color", frac=0.98)
    if same:   # edge point order is then identical: sub-pixel positions must match
        assert_close(res["edge_uv"].detach().cpu().numpy(), g["edge_uv"], 5e-3, what="edge_uv", frac=0.95)
    assert abs(float(loss) - float(g["loss"])) <= (2e-3 if same else 3e-2) * abs(float(g["loss"]))
    tol = 5e-3 if same else 6e-2     # silhouette-localised terms amplify the tracer's sub-1e-4 depth differences
    named = [("sdf." + k, p) for k, p in sdf.named_parameters()]
    for nm in ("diffuse_albedo_network", "specular_albedo_network", "specular_roughness_network", "point_light_network"):
        named += [(nm + "." + k, p) for k, p in nets[nm].named_parameters()]
    worst = 0.0
    for k, p in named:
        gr = p.grad.double().cpu()
        ref = g["gsum." + k]
        e = abs(gr.pow(2).sum().sqrt().item() - ref[2]) / max(ref[2], 1e-12)
        worst = max(worst, e)
        assert e <= tol, (k, e, same)
    print(f"[{trace_mode}] step+edges parity: identical masks={same}, worst gradient-norm error {worst:.2e}")


def test_dense_shading_matches_compacted_shading(golden):
    """dense_shading=True (shade every ray, mask afterwards: no hit-count read-back) must reproduce the reference order
    (compact the hits first) on the silhouette crop, where half of the rays miss: identical masks and dense buffers, loss
    and parameter gradients equal up to the summation order of the weight gradients."""
    g = golden("step_h256")
    outs = []
    for dense in (False, True):
        ib, sdf, nets, cam512 = build()
        cam, _, _ = cam512.crop_region(32, 32, ul_corner=tuple(int(v) for v in g["ul"]))
        rend = ib.GGXColocatedRenderer(use_cuda=True)
        loss, res = ib.stage2_step(sdf, nets, ib.RayTracer(), ib.make_render_fn(rend), cam, T(g["target"]).to(DEV),
                                   T(g["eik_points"]).to(DEV), eik_weight=0.1, dense_shading=dense)
        grads = {"sdf." + k: p.grad.clone() for k, p in sdf.named_parameters()}
        for nm in ("diffuse_albedo_network", "specular_albedo_network", "specular_roughness_network", "point_light_network"):
            grads.update({nm + "." + k: p.grad.clone() for k, p in nets[nm].named_parameters()})
        outs.append((float(loss), res, grads))
    (l0, r0, g0), (l1, r1, g1) = outs
    assert torch.equal(r0["convergent_mask"], r1["convergent_mask"])
    assert 0 < int(r0["convergent_mask"].sum()) < r0["convergent_mask"].numel()      # the crop really has misses
    for k in ("color", "diffuse_color", "specular_color", "diffuse_albedo", "specular_albedo", "specular_roughness", "normal"):
        a, b = r0[k].detach(), r1[k].detach()
        assert a.shape == b.shape, k
        assert float((a - b).abs().max()) <= 1e-6, (k, float((a - b).abs().max()))
        assert float(b[~r0["convergent_mask"]].abs().max()) == 0.0, k              # non-hit pixels are exactly zero
    assert abs(l0 - l1) <= 1e-6 * abs(l0), (l0, l1)
    worst = max(rel_l2(g1[k].cpu().numpy(), g0[k].cpu().numpy()) for k in g0)
    print(f"dense vs compact shading: loss {l1:.7f}/{l0:.7f}, worst gradient rel-L2 {worst:.2e}")
    assert worst <= 1e-4, worst


def test_graphed_step_matches_eager_step(golden):
    """GraphedStage2Step (the whole step as one CUDA-graph replay) against the eager dense-shading step: same loss and
    gradients; new inputs copied into the static buffers and in-place parameter updates are picked up by the next replay."""
    g = golden("step_h256")
    ul = tuple(int(v) for v in g["ul"])
    target, eik = T(g["target"]), T(g["eik_points"])
    Kh = torch.tensor(O.FIXTURE_K, dtype=torch.float64).reshape(4, 4).float()
    Wh = torch.tensor(O.FIXTURE_W2C, dtype=torch.float64).reshape(4, 4).float()

    def names(sdf, nets):
        out = [("sdf." + k, p) for k, p in sdf.named_parameters()]
        for nm in ("diffuse_albedo_network", "specular_albedo_network", "specular_roughness_network", "point_light_network"):
            out += [(nm + "." + k, p) for k, p in nets[nm].named_parameters()]
        return out

    def eager(sdf, nets, cam512, tgt):
        for _, p in names(sdf, nets):
            p.grad = None
        ib = __import__("iron_b200")
        cam, _, _ = ib.Camera(512, 512, Kh, Wh).crop_region(32, 32, ul_corner=ul)   # host-built, like the graph's camera
        loss, _ = ib.stage2_step(sdf, nets, ib.RayTracer(), ib.make_render_fn(ib.GGXColocatedRenderer(use_cuda=True)), cam,
                                 tgt.to(DEV), eik.to(DEV), eik_weight=0.1, dense_shading=True)
        return float(loss), {k: p.grad.clone() for k, p in names(sdf, nets)}

    ib, sdf, nets, cam512 = build()
    # the graph is captured on fresh parameters (AccumulateGrad nodes left over from an eager step on another stream would
    # invalidate the capture); the eager reference runs afterwards
    gs = ib.GraphedStage2Step(sdf, nets, ib.RayTracer(), ib.make_render_fn(ib.GGXColocatedRenderer(use_cuda=True)), Kh, Wh,
                              (32, 32), eik.shape[0], crop_ul=ul, eik_weight=0.1)
    loss = gs.step(target=target.pin_memory(), eik_points=eik.pin_memory())
    torch.cuda.synchronize()
    loss = loss.clone()
    g_graph = {k: p.grad.clone() for k, p in names(sdf, nets)}
    l_ref, g_ref = eager(sdf, nets, cam512, target)
    worst = max(rel_l2(g_graph[k].cpu().numpy(), g_ref[k].cpu().numpy()) for k in g_ref)
    print(f"graph vs eager: loss {float(loss):.7f}/{l_ref:.7f}, worst gradient rel-L2 {worst:.2e}, "
          f"{gs.kernels_per_replay} library kernels per replay")
    assert abs(float(loss) - l_ref) <= 1e-6 * abs(l_ref)
    assert worst <= 1e-5, worst          # only the atomic accumulation order of the split-K weight gradients differs
    # new target + an in-place parameter update, then replay == eager on the same state
    target2 = target * 0.5 + 0.1
    with torch.no_grad():
        sdf.lin3.weight_g.mul_(1.001)
        nets["diffuse_albedo_network"].lin1.bias.add_(0.01)
    for _, p in names(sdf, nets):
        p.grad = None
    loss2 = float(gs.step(target=target2.pin_memory()))
    g2 = {k: v.clone() for k, v in gs.grads().items()}
    l_ref2, g_ref2 = eager(sdf, nets, cam512, target2)
    g2 = {k: g2[id(p)] for k, p in names(sdf, nets)}
    assert abs(loss2 - l_ref2) <= 1e-6 * abs(l_ref2), (loss2, l_ref2)
    assert abs(loss2 - float(loss)) > 1e-4 * abs(l_ref2)            # the inputs did change the result
    errs = {k: rel_l2(g2[k].cpu().numpy(), g_ref2[k].cpu().numpy()) for k in g2}
    worst2 = max(errs.values())
    print("after the update: loss", loss2, l_ref2, "worst gradients:", sorted(errs.items(), key=lambda kv: -kv[1])[:3])
    assert worst2 <= 1e-5, worst2
    # re-sending the same camera through step() (numpy path: crop shift + fp64 inverses) reproduces the constructor's camera;
    # a different camera changes the result
    loss3 = float(gs.step(K=Kh, W2C=Wh))
    assert abs(loss3 - loss2) <= 1e-6 * abs(loss2), (loss3, loss2)
    K2 = Kh.clone()
    K2[0, 2] += 3.0
    loss4 = float(gs.step(K=K2))
    assert abs(loss4 - loss2) > 1e-4 * abs(loss2), (loss4, loss2)
    cam4, _, _ = ib.Camera(512, 512, K2, Wh).crop_region(32, 32, ul_corner=ul)
    for _, p in names(sdf, nets):
        p.grad = None
    l_ref4, _ = ib.stage2_step(sdf, nets, ib.RayTracer(), ib.make_render_fn(ib.GGXColocatedRenderer(use_cuda=True)), cam4,
                               target2.to(DEV), eik.to(DEV), eik_weight=0.1, dense_shading=True)
    assert abs(loss4 - float(l_ref4)) <= 1e-6 * abs(loss4), (loss4, float(l_ref4))
    gs.close()


def test_graphed_training_iterations_with_fused_adam_match_eager_torch_adam(golden):
    """Three complete training iterations (step + optimiser) as graph replays with FusedAdam inside the graph, against the
    eager loop with the reference's per-network torch.optim.Adam instances (lr 1e-5 / 1e-4 / 1e-2)."""
    g = golden("step_h256")
    ul = tuple(int(v) for v in g["ul"])
    target, eik = T(g["target"]), T(g["eik_points"])
    Kh = torch.tensor(O.FIXTURE_K, dtype=torch.float64).reshape(4, 4).float()
    Wh = torch.tensor(O.FIXTURE_W2C, dtype=torch.float64).reshape(4, 4).float()
    mat = ("diffuse_albedo_network", "specular_albedo_network", "specular_roughness_network")

    def groups(sdf, nets):
        return ([{"params": list(sdf.parameters()), "lr": 1e-5}] + [{"params": list(nets[nm].parameters()), "lr": 1e-4} for nm in mat]
                + [{"params": list(nets["point_light_network"].parameters()), "lr": 1e-2}])

    ib, sdf, nets, _ = build()
    opt = ib.FusedAdam(groups(sdf, nets))
    rf = ib.make_render_fn(ib.GGXColocatedRenderer(use_cuda=True))
    start = {k: v.clone() for k, v in sdf.state_dict().items()}
    gs = ib.GraphedStage2Step(sdf, nets, ib.RayTracer(), rf, Kh, Wh, (32, 32), eik.shape[0], crop_ul=ul, eik_weight=0.1,
                              optimizer=opt)
    for k, v in sdf.state_dict().items():
        assert torch.equal(v, start[k]), f"warm-up / capture moved {k}"
    losses = []
    for _ in range(3):
        losses.append(float(gs.step(target=target.pin_memory(), eik_points=eik.pin_memory())))
    torch.cuda.synchronize()
    assert opt.steps_taken == 3

    ib2, sdf2, nets2, _ = build()
    opts = [torch.optim.Adam(gr["params"], lr=gr["lr"]) for gr in groups(sdf2, nets2)]
    cam, _, _ = ib2.Camera(512, 512, Kh, Wh).crop_region(32, 32, ul_corner=ul)
    ref_losses = []
    for _ in range(3):
        for o in opts:
            o.zero_grad()
        loss, _ = ib2.stage2_step(sdf2, nets2, ib2.RayTracer(), rf, cam, target.to(DEV), eik.to(DEV), eik_weight=0.1,
                                  dense_shading=True)
        ref_losses.append(float(loss))
        for o in opts:
            o.step()
    print("graph + FusedAdam losses", losses, "eager + torch.optim.Adam losses", ref_losses)
    assert losses[0] != losses[2]
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) <= 1e-5 * abs(b), (losses, ref_losses)
    worst = 0.0
    for (k, p), (_, q) in zip(list(sdf.named_parameters()) + [x for nm in mat for x in nets[nm].named_parameters()],
                              list(sdf2.named_parameters()) + [x for nm in mat for x in nets2[nm].named_parameters()]):
        worst = max(worst, rel_l2(p.detach().cpu().numpy(), q.detach().cpu().numpy()))
    # Adam's first steps move every element by ~lr * sign(g): elements whose gradient is ~0 (zero-initialised biases) turn the
    # atomic-order noise of the weight gradients into O(lr) differences, so the bound is loose in relative terms
    assert worst <= 5e-4, worst
    gs.close()


def test_graphed_step_bench_configuration_matches_eager():
    """The bench workload itself (BASELINE configs[1]: H = 512, 64 x 64 centre crop, 2,048 eikonal points): the graph replay
    with its parallel streams (eikonal branch, three material MLPs) against the eager single-stream step -- loss and every
    parameter gradient -- over several replays."""
    import iron_b200 as ib
    torch.manual_seed(0)
    nets = ib.init_rendering_network_dict("ggx")
    torch.manual_seed(0)
    sdf = ib.SDFNetwork(d_in=3, d_out=257, d_hidden=512, n_layers=8, skip_in=[4], multires=6, bias=0.5, scale=1.0,
                        geometric_init=True, weight_norm=True).to(DEV)
    nets["point_light_network"].set_light(32.0)
    Kh = torch.tensor(O.FIXTURE_K, dtype=torch.float64).reshape(4, 4).float()
    Wh = torch.tensor(O.FIXTURE_W2C, dtype=torch.float64).reshape(4, 4).float()
    S, ul = 64, (224, 224)
    gen = torch.Generator().manual_seed(11)
    target = (torch.rand(S, S, 3, generator=gen) * 0.5)
    eik = torch.empty(S * S // 2, 3).uniform_(-1.0, 1.0, generator=gen)
    rf = ib.make_render_fn(ib.GGXColocatedRenderer(use_cuda=True))
    params = [("sdf." + k, p) for k, p in sdf.named_parameters()]
    for nm in ("diffuse_albedo_network", "specular_albedo_network", "specular_roughness_network", "point_light_network"):
        params += [(nm + "." + k, p) for k, p in nets[nm].named_parameters()]
    gs = ib.GraphedStage2Step(sdf, nets, ib.RayTracer(), rf, Kh, Wh, (S, S), eik.shape[0], crop_ul=ul)
    assert gs._eik_stream is not None and gs._mat_streams is not None       # the overlapped configuration is what is tested
    graph_runs = []
    for _ in range(5):
        loss = gs.step(target=target.pin_memory(), eik_points=eik.pin_memory())
        torch.cuda.synchronize()
        graph_runs.append((float(loss), {k: p.grad.clone() for k, p in params}))
    cam, _, _ = ib.Camera(512, 512, Kh, Wh).crop_region(S, S, ul_corner=ul)
    for _, p in params:
        p.grad = None
    l_ref, _ = ib.stage2_step(sdf, nets, ib.RayTracer(), rf, cam, target.to(DEV), eik.to(DEV), dense_shading=True)
    g_ref = {k: p.grad.clone() for k, p in params}
    for lg, gg in graph_runs:
        assert abs(lg - float(l_ref)) <= 1e-6 * abs(float(l_ref)), (lg, float(l_ref))
        worst = max(rel_l2(gg[k].cpu().numpy(), g_ref[k].cpu().numpy()) for k in g_ref)
        assert worst <= 1e-5, worst
    print(f"bench configuration: loss {graph_runs[0][0]:.7f} / {float(l_ref):.7f}, 5 replays, gradients within 1e-5")
    gs.close()
    # the data-parallel variant: the graph's tail packs every gradient (x grad_scale = 1 / world) into ONE flat bucket and
    # .grad becomes a view of it, so the exchange is a single in-place all-reduce (parallel.allreduce_flat)
    gf = ib.GraphedStage2Step(sdf, nets, ib.RayTracer(), rf, Kh, Wh, (S, S), eik.shape[0], crop_ul=ul, flat_grads=True,
                              grad_scale=0.5)
    for _ in range(2):
        gf.step(target=target.pin_memory(), eik_points=eik.pin_memory())
    torch.cuda.synchronize()
    lo, hi = gf.flat_grad.data_ptr(), gf.flat_grad.data_ptr() + gf.flat_grad.numel() * 4
    for k, p in params:
        assert lo <= p.grad.data_ptr() < hi, k                               # a view of the bucket
        assert rel_l2(p.grad.cpu().numpy() * 2.0, g_ref[k].cpu().numpy()) <= 1e-5, k
    assert gf.flat_grad.numel() >= sum(p.numel() for _, p in params)
    gf.close()


def test_eager_steps_with_fused_adam_match_torch_adam(golden):
    """The advertised drop-in: an EAGER training loop with iron_b200.FusedAdam against the same loop with the reference's
    torch.optim.Adam instances.  FusedAdam writes the parameters through raw pointers; unless it bumps their version counters
    the weight-norm fold cache (keyed on p._version) keeps serving the initial weights and the losses never move."""
    g = golden("step_h256")
    ul = tuple(int(v) for v in g["ul"])
    target, eik = T(g["target"]).to(DEV), T(g["eik_points"]).to(DEV)
    mat = ("diffuse_albedo_network", "specular_albedo_network", "specular_roughness_network")

    def groups(sdf, nets):
        return ([{"params": list(sdf.parameters()), "lr": 1e-4}] + [{"params": list(nets[nm].parameters()), "lr": 1e-3} for nm in mat]
                + [{"params": list(nets["point_light_network"].parameters()), "lr": 1e-2}])

    def run(make_opts):
        ib, sdf, nets, cam512 = build()
        cam, _, _ = cam512.crop_region(32, 32, ul_corner=ul)
        opts = make_opts(ib, groups(sdf, nets))
        rf = ib.make_render_fn(ib.GGXColocatedRenderer(use_cuda=True))
        losses = []
        for _ in range(4):
            for o in opts:
                o.zero_grad()
            loss, _ = ib.stage2_step(sdf, nets, ib.RayTracer(), rf, cam, target, eik, eik_weight=0.1, dense_shading=True)
            losses.append(float(loss))
            for o in opts:
                o.step()
        probe = sdf.sdf(eik[:64]).detach().cpu().numpy()        # an eager forward AFTER the last optimiser step
        return losses, probe

    fused, probe_f = run(lambda ib, gr: [ib.FusedAdam(gr)])
    ref, probe_r = run(lambda ib, gr: [torch.optim.Adam(x["params"], lr=x["lr"]) for x in gr])
    print("eager + FusedAdam", fused, "eager + torch.optim.Adam", ref)
    assert abs(ref[0] - ref[3]) > 1e-4 * abs(ref[0]), "the reference loop itself did not move"
    for a, b in zip(fused, ref):
        assert abs(a - b) <= 2e-5 * abs(b), (fused, ref)
    assert np.abs(probe_f - probe_r).max() <= 1e-5


def test_graph_with_fused_adam_sees_lr_changes_and_invalidates_eager_folds(golden):
    """A captured optimiser step never runs its Python again: GraphedStage2Step.step() has to refresh the pinned
    pointer / hyper-parameter table (lr schedules) before the replay and bump the parameter versions after it."""
    g = golden("step_h256")
    ul = tuple(int(v) for v in g["ul"])
    target, eik = T(g["target"]), T(g["eik_points"])
    Kh = torch.tensor(O.FIXTURE_K, dtype=torch.float64).reshape(4, 4).float()
    Wh = torch.tensor(O.FIXTURE_W2C, dtype=torch.float64).reshape(4, 4).float()
    ib, sdf, nets, _ = build()
    opt = ib.FusedAdam([{"params": list(sdf.parameters()), "lr": 1e-4}])
    rf = ib.make_render_fn(ib.GGXColocatedRenderer(use_cuda=True))
    gs = ib.GraphedStage2Step(sdf, nets, ib.RayTracer(), rf, Kh, Wh, (32, 32), eik.shape[0], crop_ul=ul, eik_weight=0.1,
                              optimizer=opt)
    x = eik[:128].to(DEV)
    f0 = sdf.sdf(x).detach().clone()
    w0 = sdf.lin3.weight_v.detach().clone()
    gs.step(target=target.pin_memory(), eik_points=eik.pin_memory())
    torch.cuda.synchronize()
    d1 = float((sdf.lin3.weight_v.detach() - w0).abs().max())
    f1 = sdf.sdf(x).detach().clone()                               # eager call after a replay: must see the new weights
    assert d1 > 0 and float((f1 - f0).abs().max()) > 0, "eager forward after a replay still used the pre-step fold"
    opt.param_groups[0]["lr"] = 0.0                                # a scheduler sets lr: the next replay must not move anything
    w1 = sdf.lin3.weight_v.detach().clone()
    gs.step()
    torch.cuda.synchronize()
    assert float((sdf.lin3.weight_v.detach() - w1).abs().max()) == 0.0, "the replay ignored the new learning rate"
    # checkpoint / resume: the moments keep their storage (the graph holds pointers to it), values and step count come back
    import copy
    sd = copy.deepcopy(opt.state_dict())
    assert sd["ironb_step"] == 2
    m_ptr = opt.state[sdf.lin3.weight_v]["exp_avg"].data_ptr()
    m_val = opt.state[sdf.lin3.weight_v]["exp_avg"].clone()
    opt.state[sdf.lin3.weight_v]["exp_avg"].zero_()
    opt._step_dev.zero_()
    opt.load_state_dict(sd)
    assert opt.state[sdf.lin3.weight_v]["exp_avg"].data_ptr() == m_ptr
    assert torch.equal(opt.state[sdf.lin3.weight_v]["exp_avg"], m_val) and opt.steps_taken == 2
    gs.close()


H512_CROPS = {"centre": (224, 224), "silhouette": (342, 224)}


@pytest.mark.parametrize("crop", list(H512_CROPS))
def test_step_bench_configuration_against_oracle(crop, trace_mode):
    """The configuration bench.py times (BASELINE configs[1]: H = 512, 64 x 64 crop, 2,048 eikonal points, seed-0 weights),
    full step against the CPU oracle's stage2_step (the restated reference, pinned by tests/golden): hit mask, distance,
    normal, RGB, loss and every parameter-gradient tensor at the BASELINE tolerances.  `silhouette` straddles the object's
    outline (r ~ 0.287 -> ~118 px from the image centre), where rays graze and the sampler / bisection do the work."""
    import iron_b200 as ib
    S, ul = 64, H512_CROPS[crop]
    torch.manual_seed(0)
    nets = ib.init_rendering_network_dict("ggx")
    torch.manual_seed(0)
    sdf = ib.SDFNetwork(d_in=3, d_out=257, d_hidden=512, n_layers=8, skip_in=[4], multires=6, bias=0.5, scale=1.0,
                        geometric_init=True, weight_norm=True)
    sdf_p = {k: v.detach().clone().requires_grad_(True) for k, v in sdf.state_dict().items()}
    mat = ("diffuse_albedo_network", "specular_albedo_network", "specular_roughness_network")
    mats_p = {nm: {k: v.detach().cpu().clone().requires_grad_(True) for k, v in nets[nm].state_dict().items()} for nm in mat}
    light_p = torch.tensor(32.0, requires_grad=True)
    sdf = sdf.to(DEV)
    nets["point_light_network"].set_light(32.0)
    target = torch.rand(S, S, 3, generator=torch.Generator().manual_seed(11)) * 0.5
    eik = torch.empty(S * S // 2, 3).uniform_(-1.0, 1.0, generator=torch.Generator().manual_seed(12))
    K = torch.tensor(O.FIXTURE_K, dtype=torch.float64).reshape(4, 4).float()
    W2C = torch.tensor(O.FIXTURE_W2C, dtype=torch.float64).reshape(4, 4).float()
    cam, _, _ = ib.Camera(512, 512, K.to(DEV), W2C.to(DEV)).crop_region(S, S, ul_corner=ul)
    rf = ib.make_render_fn(ib.GGXColocatedRenderer(use_cuda=True))
    loss, res = ib.stage2_step(sdf, nets, ib.RayTracer(), rf, cam, target.to(DEV), eik.to(DEV), dense_shading=True)
    torch.cuda.synchronize()
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    oloss, ores = O.stage2_step(sdf_p, mats_p, light_p, O.OCamera.fixture().crop(S, S, ul), target, eik.clone())
    m, mr = res["convergent_mask"].cpu().numpy(), ores["convergent_mask"].numpy()
    agree = (m == mr).mean()
    both = m & mr
    assert both.sum() > 500, f"crop {crop}: only {both.sum()} common hits"
    if crop == "silhouette":
        assert (~mr).sum() > 500, "the silhouette crop must contain misses"
    nerr = np.abs(res["normal"].detach().cpu().numpy()[both] - ores["normal"].detach().numpy()[both]).max(axis=-1)
    derr = np.abs(res["distance"].cpu().numpy()[both] - ores["distance"].numpy()[both])
    cerr = np.abs(res["color"].detach().cpu().numpy()[both] - ores["color"].detach().numpy()[both]).max(axis=-1)
    same_mask = bool((m == mr).all())
    print(f"[H512 {crop} / {trace_mode}] hits {int(m.sum())}/{int(mr.sum())} mask-agree {agree:.5f}  normals<=1e-4: "
          f"{100 * (nerr <= 1e-4).mean():.2f}% (max {nerr.max():.2e})  dist<=1e-4: {100 * (derr <= 1e-4).mean():.2f}% "
          f"(max {derr.max():.2e})  rgb max {cerr.max():.2e}  loss {float(loss):.7f} / {float(oloss):.7f}")
    assert agree >= 0.9995, f"hit mask agreement {agree:.5f}"                     # <= 2 of 4,096 rays (pooled check: test_trace_gpu)
    assert (derr <= 1e-4).mean() >= 0.995 and derr.max() <= 2.0 / 127, "distance"
    need = normal_need("h512_" + crop)
    assert (nerr <= TOL_NORMAL).mean() >= need, f"normals within 1e-4: {(nerr <= 1e-4).mean():.4f}, need {need:.4f}"
    assert nerr.max() <= 2e-3
    assert (cerr <= TOL_RGB).mean() >= 0.995, "rgb"
    assert abs(float(loss) - float(oloss)) <= (1e-3 if same_mask else 2e-2) * abs(float(oloss)), (float(loss), float(oloss))
    tol = TOL_GRAD_REL if same_mask else 5e-2
    named = [("sdf." + k, p, sdf_p[k]) for k, p in sdf.named_parameters()]
    for nm in mat:
        named += [(nm + "." + k, p, mats_p[nm][k]) for k, p in nets[nm].named_parameters()]
    worst = 0.0
    for k, p, q in named:
        assert p.grad is not None and q.grad is not None, k
        r = rel_l2(p.grad.cpu().numpy(), q.grad.numpy())
        worst = max(worst, r)
        assert r <= tol, (k, r, "same_mask", same_mask)
    gl = float(nets["point_light_network"].light.grad) if hasattr(nets["point_light_network"], "light") else None
    if gl is not None:
        assert abs(gl - float(light_p.grad)) <= tol * abs(float(light_p.grad)), (gl, float(light_p.grad))
    print(f"[H512 {crop} / {trace_mode}] same_mask={same_mask} worst relative gradient error {worst:.2e}")


def test_step_with_the_reference_loss_golden(golden, trace_mode):
    """The step with the loss the reference trains with (PyramidL2 + SSIM + roughness range + eikonal,
    render_surface.py:594-613; stage2_step(image_loss="reference")) against a step of the real reference modules
    (tests/golden/step_refloss_h256.npz), eager and as a graph replay."""
    g = golden("step_refloss_h256")
    ib, sdf, nets, cam512 = build()
    with torch.no_grad():
        nets["specular_roughness_network"].lin4.bias.add_(float(g["rough_bias_shift"]))
    ul = tuple(int(v) for v in g["ul"])
    cam, _, _ = cam512.crop_region(32, 32, ul_corner=ul)
    rf = ib.make_render_fn(ib.GGXColocatedRenderer(use_cuda=True))
    target, eik = T(g["target"]), T(g["eik_points"])
    named = [("sdf." + k, p) for k, p in sdf.named_parameters()]
    for nm in ("diffuse_albedo_network", "specular_albedo_network", "specular_roughness_network", "point_light_network"):
        named += [(nm + "." + k, p) for k, p in nets[nm].named_parameters()]

    def check(loss, res, what):
        m = res["convergent_mask"].cpu().numpy()
        same_mask = bool((m == g["mask"]).all())
        assert (m == g["mask"]).mean() >= 0.998
        assert abs(float(loss) - float(g["loss"])) <= (1e-3 if same_mask else 2e-2) * abs(float(g["loss"])), (what, float(loss), float(g["loss"]))
        tol = TOL_GRAD_REL if same_mask else 5e-2
        worst = 0.0
        for k, p in named:
            gr = p.grad.double().cpu()
            ref = g["gsum." + k]
            e = abs(gr.pow(2).sum().sqrt().item() - ref[2]) / max(ref[2], 1e-12)
            worst = max(worst, e)
            assert e <= tol, (what, k, gr.pow(2).sum().sqrt().item(), ref[2])
            if "g." + k in g:
                r = rel_l2(gr.numpy(), g["g." + k])
                worst = max(worst, r)
                assert r <= 2 * tol, (what, k, r)
        print(f"[{trace_mode}] reference-loss step ({what}): loss {float(loss):.6f} / {float(g['loss']):.6f}, same_mask={same_mask}, "
              f"worst relative gradient error {worst:.2e}")

    loss, res = ib.stage2_step(sdf, nets, ib.RayTracer(), rf, cam, target.to(DEV), eik.to(DEV), eik_weight=0.1,
                               image_loss="reference")
    check(loss, res, "eager, compacted shading")
    if not trace_mode.startswith("batched-tcgen05"):
        return          # the graph replay is built on the default tracer (the fused FFMA tracer is a diagnostic mode)
    # fresh modules for the graph: it has to be captured before any eager backward through the same parameters
    ib, sdf, nets, cam512 = build()
    with torch.no_grad():
        nets["specular_roughness_network"].lin4.bias.add_(float(g["rough_bias_shift"]))
    named[:] = [("sdf." + k, p) for k, p in sdf.named_parameters()]
    for nm in ("diffuse_albedo_network", "specular_albedo_network", "specular_roughness_network", "point_light_network"):
        named += [(nm + "." + k, p) for k, p in nets[nm].named_parameters()]
    Kh = torch.tensor(O.FIXTURE_K, dtype=torch.float64).reshape(4, 4).float()
    Wh = torch.tensor(O.FIXTURE_W2C, dtype=torch.float64).reshape(4, 4).float()
    gs = ib.GraphedStage2Step(sdf, nets, ib.RayTracer(), rf, Kh, Wh, (32, 32), eik.shape[0], crop_ul=ul, eik_weight=0.1,
                              image_loss="reference")
    for _ in range(2):
        loss = gs.step(target=target.pin_memory(), eik_points=eik.pin_memory())
        torch.cuda.synchronize()
    check(loss, gs.results, "graph replay")
    gs.close()
