import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
        return cache[name]

    return load


GEMM_MODES = {"ffma": 0, "tcgen05-tf32": 1, "tcgen05": 2}


@pytest.fixture(params=list(GEMM_MODES))
def gemm_mode(request):
    """Runs a GPU test under every GEMM arithmetic of the differentiable part: fp32 FFMA tiles (tight value
    tolerances), the tcgen05 3xTF32 kernel everywhere ("tcgen05-tf32"; the tensor core accumulates with truncation, so
    forward values carry ~1e-5 relative error -- still far inside the BASELINE tolerances, which the tests also assert),
    and the default "tcgen05": fp16x2 pre-split operands for the forward-type products (de-biased accumulators), 3xTF32
    for products with gradient operands."""
    from iron_b200 import _lib
    lib = _lib.load()
    prev = lib.ironb_set_gemm_mode(GEMM_MODES[request.param])
    yield request.param
    lib.ironb_set_gemm_mode(prev)


TRACE_MODES = {"fused-ffma": 0, "batched-tcgen05": 2, "batched-tcgen05-wide": 2}


@pytest.fixture(params=list(TRACE_MODES))
def trace_mode(request):
    """Both tracer implementations: the fused persistent fp32-FFMA kernels (exact fp32 association) and the batched
    tcgen05 rounds with fp16x2-split operands and two tiles in flight (default; both cluster shapes of its MLP kernel).
    (The 3xTF32 predecessor of the latter was retired in round 2.)"""
    from iron_b200 import _lib
    lib = _lib.load()
    prev = lib.ironb_set_trace_mode(TRACE_MODES[request.param])
    # the fused MLP kernel has two shapes (csrc/mlp_h16.cu): automatic selection takes the 64-column one for the small ray
    # counts of these tests; "-wide" forces the 128-column one, which large patches use
    prev_rn = lib.ironb_set_mlp_rn(128 if request.param.endswith("-wide") else 0)
    yield request.param
    lib.ironb_set_mlp_rn(prev_rn)
    lib.ironb_set_trace_mode(prev)
