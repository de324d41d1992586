"""CPU-side checks: the C-ABI library builds/loads, exports every symbol include/iron_b200.h declares, and the
host-only layout functions agree with the reference's layer shapes.  No kernels are launched."""
import ctypes as C
import os
import re

import pytest

from util import ROOT


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "iron_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ironb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from iron_b200 import _lib
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/iron_b200.h but not exported"
    assert set(syms) == set(_lib.EXPORTS), set(syms) ^ set(_lib.EXPORTS)
    assert lib.ironb_version() >= 100


def test_sdf_layout_matches_reference_shapes():
    from iron_b200 import _lib
    lib = _lib.load()
    for H in (256, 512):
        L = _lib.MlpLayout()
        assert lib.ironb_sdf_layout(3, 257, H, 8, 4, 6, 1.0, 100.0, C.byref(L)) == 0
        assert L.n_lin == 9 and L.pe_dim == 39
        assert list(L.in_dim)[:9] == [39] + [H] * 8                      # models/fields.py:26-45
        assert list(L.out_dim)[:9] == [H, H, H, H - 39, H, H, H, H, 257]
        assert list(L.in_pad)[:9] == [40] + [H] * 8
        assert list(L.out_pad)[:9] == [H] * 8 + [264]
        n_params = sum(L.in_dim[l] * L.out_dim[l] + 2 * L.out_dim[l] for l in range(9))
        assert n_params == {256: 529076, 512: 1975220}[H]
    L = _lib.MlpLayout()
    assert lib.ironb_sdf_layout(2, 257, 256, 8, 4, 6, 1.0, 100.0, C.byref(L)) != 0
    assert b"d_in" in lib.ironb_last_error()


def test_matnet_layout_and_input_width():
    from iron_b200 import _lib
    lib = _lib.load()
    for mode, mr, mrv, want in ((0, 0, 4, 289), (1, 6, 0, 298)):          # models/network_conf.py:72-120
        cfg = _lib.MatnetCfg(mode, mr, mrv, 256, 0, 0.0, 1.0, 1.0)
        assert lib.ironb_matnet_in_dim(C.byref(cfg)) == want
        L = _lib.MlpLayout()
        assert lib.ironb_matnet_layout(want, 3, 256, 4, C.byref(L)) == 0
        assert L.n_lin == 5 and L.in_dim[0] == want and L.in_pad[0] % 8 == 0 and L.out_pad[4] == 8


def test_modules_refuse_cpu_tensors():
    import torch
    import iron_b200
    net = iron_b200.SDFNetwork(3, 17, 64, 8, skip_in=[4], multires=6)
    with pytest.raises(RuntimeError, match="CUDA"):
        net(torch.zeros(4, 3))
    r = iron_b200.GGXColocatedRenderer()
    with pytest.raises(RuntimeError, match="CUDA"):
        r(torch.tensor(1.0), torch.ones(2, 1), torch.ones(2, 3), torch.ones(2, 3),
          {"diffuse_albedo": torch.ones(2, 3), "specular_albedo": torch.ones(2, 3), "specular_roughness": torch.ones(2, 1)})


def test_state_dict_keys_and_seeded_init_match_reference(golden):
    """Same ctor RNG consumption as models/fields.py: torch.manual_seed(0) reproduces the reference weights."""
    import numpy as np
    import torch
    import iron_b200
    g = golden("sdf_seeded")
    for H in (256, 512):
        torch.manual_seed(0)
        net = iron_b200.SDFNetwork(d_in=3, d_out=257, d_hidden=H, n_layers=8, skip_in=[4], multires=6, bias=0.5,
                                   scale=1.0, geometric_init=True, weight_norm=True)
        sd = net.state_dict()
        assert sorted(sd) == sorted(f"lin{l}.{k}" for l in range(9) for k in ("weight_g", "weight_v", "bias"))
        for k, v in sd.items():
            ref = g[f"h{H}.sum.{k}"]
            got = np.array([v.double().sum().item(), v.double().abs().sum().item()])
            assert np.allclose(got, ref, rtol=1e-6, atol=1e-6), (H, k, got, ref)
