"""fill_holes (row f-4): 3x3 depth closing kernel and the mask / depth / points update of raytrace_camera, against the
oracle's restatement (models/raytracer.py:552-564), and against tests/golden/morph_cv2.npz: OpenCV's closing / Sobel of the
same depth maps (oracle/make_golden_cv2.py), the independent pin for the two kornia calls (models/raytracer.py:557, 569)."""
import numpy as np
import pytest
import torch

from oracle import iron_oracle as O
from util import assert_close, perturb

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("H,W", [(1, 1), (3, 5), (64, 64), (127, 33)])
def test_depth_closing_kernel(H, W):
    from iron_b200.raytracer import depth_closing
    g = torch.Generator().manual_seed(H * 100 + W)
    d = torch.rand(H, W, generator=g) + 1.0
    d[torch.rand(H, W, generator=g) < 0.2] = 0.0            # holes
    got = depth_closing(d.to(DEV)).cpu()
    assert torch.equal(got, O.morph_closing3(d))


def test_closing_and_sobel_kernels_match_opencv(golden):
    """ironb_depth_closing bit-exact, ironb_sobel_depth within 5e-7, against cv2.morphologyEx(MORPH_CLOSE) / cv2.Sobel."""
    from iron_b200.raytracer import depth_closing, sobel_depth
    g = golden("morph_cv2")
    for i in range(int(g["n"])):
        d = torch.from_numpy(g[f"depth{i}"]).to(DEV)
        assert np.array_equal(depth_closing(d).cpu().numpy(), g[f"closing{i}"]), f"closing, image {i}"
        assert_close(sobel_depth(d).cpu().numpy(), g[f"sobel{i}"], 5e-7, what=f"sobel, image {i}")


def test_raytrace_camera_fill_holes(trace_mode):
    import iron_b200 as ib
    torch.manual_seed(0)
    sdf = ib.SDFNetwork(d_in=3, d_out=257, d_hidden=256, n_layers=8, skip_in=[4], multires=6, bias=0.5, scale=1.0,
                        geometric_init=True, weight_norm=True)
    perturb(sdf, 0.005, seed=1)
    sdf = sdf.to(DEV)
    K = torch.tensor(O.FIXTURE_K, dtype=torch.float64).reshape(4, 4).float().to(DEV)
    W2C = torch.tensor(O.FIXTURE_W2C, dtype=torch.float64).reshape(4, 4).float().to(DEV)
    cam, _, _ = ib.Camera(512, 512, K, W2C).crop_region(48, 48, ul_corner=(426, 232))
    rt = ib.RayTracer()
    plain = ib.raytrace_camera(cam, sdf, rt, max_num_rays=50000, fill_holes=False)
    # punch holes into the traced result the way a failed ray would look, then run the hole filling on it
    holes = torch.zeros(48, 48, dtype=torch.bool)
    holes[10, 5] = holes[20, 7] = holes[21, 8] = holes[30, 3] = True
    holes = holes.to(DEV) & plain["convergent_mask"]
    assert int(holes.sum()) >= 2
    inp = {k: v.clone() for k, v in plain.items()}
    inp["convergent_mask"] = plain["convergent_mask"] & ~holes
    ref = O.fill_holes({k: v.cpu() for k, v in inp.items()})
    # device path: same update through the module's code
    import iron_b200.raytracer as R
    orig = R.raytrace_pixels
    R.raytrace_pixels = lambda *a, **k: {kk: vv.clone() for kk, vv in inp.items()}
    try:
        got = ib.raytrace_camera(cam, sdf, rt, max_num_rays=50000, fill_holes=True)
    finally:
        R.raytrace_pixels = orig
    assert torch.equal(got["convergent_mask"].cpu(), ref["convergent_mask"])
    assert bool(got["convergent_mask"][holes].all())              # isolated holes are closed
    for k in ("depth", "distance", "points"):
        assert_close(got[k].cpu().numpy(), ref[k].numpy(), 1e-6, what=k)
    # no holes -> nothing changes (the reference only re-derives distance / points when something was filled)
    got2 = ib.raytrace_camera(cam, sdf, rt, max_num_rays=50000, fill_holes=True)
    ref2 = O.fill_holes({k: v.cpu() for k, v in plain.items()})
    assert torch.equal(got2["convergent_mask"].cpu(), ref2["convergent_mask"])
    assert_close(got2["points"].cpu().numpy(), ref2["points"].numpy(), 1e-6, what="points (no update)")
