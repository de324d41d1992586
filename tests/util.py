"""Shared helpers for the parity tests (CPU oracle vs CUDA path)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# BASELINE.json tolerances (north_star): hit mask >= 99.99 %, depth / normal 1e-4, RGB 1e-3, parameter grads 1e-3 relative
TOL_DEPTH = 1e-4
TOL_NORMAL = 1e-4
TOL_RGB = 1e-3
TOL_GRAD_REL = 1e-3


def T(a):
    return torch.from_numpy(np.asarray(a).copy())


def err_stats(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    e = np.abs(a - b)
    return e.max() if e.size else 0.0, e


def assert_close(a, b, atol, rtol=0.0, what="", frac=1.0):
    """|a-b| <= atol + rtol*|b| for at least `frac` of the entries (frac<1 states a quantile tolerance)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if a.size == 0:
        return
    e = np.abs(a - b)
    tol = atol + rtol * np.abs(b)
    ok = (e <= tol)
    got = ok.mean()
    assert got >= frac, (f"{what}: {100 * got:.4f}% within tol (need {100 * frac:.4f}%), max err {e.max():.3e} at "
                         f"{np.unravel_index(np.argmax(e - tol), e.shape)} (ref {b.flat[np.argmax(e - tol)]:.6g}), "
                         f"atol {atol:.1e} rtol {rtol:.1e}")


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def perturb(sdf_net, sigma, seed=1):
    """'shape' variant of oracle/make_golden.py: Gaussian noise on lin1..lin7.weight_v (CPU generator)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for l in range(1, 8):
            v = getattr(sdf_net, f"lin{l}").weight_v
            v.add_((torch.randn(v.shape, generator=g) * sigma).to(v.device))


def params_np(module):
    return {k: v.detach().cpu() for k, v in module.state_dict().items()}


def oracle_params(module):
    """state_dict of an iron_b200 module -> the oracle's parameter dict (CPU tensors)."""
    return {k: v.detach().cpu().clone() for k, v in module.state_dict().items()}
