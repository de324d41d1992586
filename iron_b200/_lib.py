"""ctypes binding of libiron_b200.so (include/iron_b200.h).  Loading fails loudly: no library, no product."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch

from . import build as _build

MAX_LIN = 12


class MlpLayout(C.Structure):
    _fields_ = [
        ("n_lin", C.c_int32), ("kind", C.c_int32), ("d_in", C.c_int32), ("multires", C.c_int32),
        ("pe_dim", C.c_int32), ("skip_layer", C.c_int32), ("d_hidden", C.c_int32), ("d_out", C.c_int32),
        ("scale", C.c_float), ("beta", C.c_float),
        ("in_dim", C.c_int32 * MAX_LIN), ("out_dim", C.c_int32 * MAX_LIN),
        ("in_pad", C.c_int32 * MAX_LIN), ("out_pad", C.c_int32 * MAX_LIN),
        ("off_w", C.c_int64 * MAX_LIN), ("off_wt", C.c_int64 * MAX_LIN), ("off_b", C.c_int64 * MAX_LIN),
        ("packed_floats", C.c_int64),
        ("off_h16", C.c_int64 * MAX_LIN), ("off_h16t", C.c_int64 * MAX_LIN), ("packed_total_floats", C.c_int64),
    ]


class MatnetCfg(C.Structure):
    _fields_ = [
        ("mode", C.c_int32), ("multires", C.c_int32), ("multires_view", C.c_int32), ("d_feature", C.c_int32),
        ("squeeze", C.c_int32), ("out_bias", C.c_float), ("out_scale", C.c_float), ("squeeze_scale", C.c_float),
    ]


_P = C.c_void_p
_I64 = C.c_int64
_INT = C.c_int
_F = C.c_float
_LAY = C.POINTER(MlpLayout)
_CFG = C.POINTER(MatnetCfg)
_PP = C.POINTER(C.c_void_p)
_PI64 = C.POINTER(C.c_int64)

_PROTOS = {
    "ironb_last_error": (C.c_char_p, []),
    "ironb_version": (_INT, []),
    "ironb_launch_count": (_I64, []),
    "ironb_set_gemm_mode": (_INT, [_INT]),
    "ironb_set_trace_mode": (_INT, [_INT]),
    "ironb_set_mlp_debias": (_F, [_F]),
    "ironb_set_mlp_rn": (_INT, [_INT]),
    "ironb_debug_hang_buffer": (_INT, [_P]),
    "ironb_gemm_nt": (_INT, [_P, _INT, _P, _INT, _INT, _INT, _INT, _P, _INT, _INT, _P]),
    "ironb_gemm_nt_h16": (_INT, [_P, _P, _INT, _P, _P, _INT, _INT, _INT, _INT, _P, _INT, _P]),
    "ironb_gemm_tn_scratch_bytes": (_I64, [_INT, _INT, _INT]),
    "ironb_gemm_tn": (_INT, [_P, _INT, _P, _INT, _INT, _INT, _INT, _P, _INT, _INT, _P, _P]),
    "ironb_sdf_layout": (_INT, [_INT, _INT, _INT, _INT, _INT, _INT, _F, _F, _LAY]),
    "ironb_matnet_layout": (_INT, [_INT, _INT, _INT, _INT, _LAY]),
    "ironb_matnet_layout_skip": (_INT, [_INT, _INT, _INT, _INT, _INT, _LAY]),
    "ironb_mlp_fold": (_INT, [_LAY, _PP, _PP, _PP, _P, _P]),
    "ironb_mlp_fold_bwd": (_INT, [_LAY, _PP, _PP, _P, _PP, _PP, _PP, _P]),
    "ironb_sdf_getall_workspace_bytes": (_I64, [_LAY, _I64, _INT, _INT]),
    "ironb_sdf_getall_fwd": (_INT, [_LAY, _P, _P, _I64, _P, _P, _P, _INT, _P, _I64, _P]),
    "ironb_sdf_getall_bwd": (_INT, [_LAY, _P, _P, _I64, _P, _P, _P, _P, _I64, _P, _P, _P]),
    "ironb_matnet_in_dim": (_INT, [_CFG]),
    "ironb_matnet_workspace_bytes": (_I64, [_LAY, _I64]),
    "ironb_matnet_fwd": (_INT, [_LAY, _CFG, _P, _P, _P, _P, _P, _I64, _P, _P, _I64, _P]),
    "ironb_matnet_bwd": (_INT, [_LAY, _CFG, _P, _I64, _P, _P, _P, _I64, _P, _P, _P, _P, _P, _P, _P]),
    "ironb_ggx_fwd": (_INT, [_P] * 9 + [_I64] + [_P] * 4),
    "ironb_ggx_bwd": (_INT, [_P] * 9 + [_I64] + [_P] * 11),
    "ironb_composite_fwd": (_INT, [_P] * 12 + [_I64] + [_P] * 5),
    "ironb_composite_bwd": (_INT, [_P] * 12 + [_I64] + [_P] * 14),
    "ironb_camera_rays": (_INT, [_P, _I64, _P, _P, _P, _F, _P, _P, _P, _P, _P, _P, _P]),
    "ironb_intersect_sphere": (_INT, [_P, _P, _I64, _F, _P, _P, _P, _P]),
    "ironb_trace_workspace_bytes": (_I64, [_LAY, _I64]),
    "ironb_trace": (_INT, [_LAY, _P, _P, _P, _P, _P, _P, _I64, _F, _INT, _INT, _P, _P, _P, _P, _P, _P, _P, _I64, _P]),
    "ironb_compact_mask": (_INT, [_P, _I64, _P, _P, _P]),
    "ironb_gather_rows": (_INT, [_P, _P, _I64, _INT, _P, _P]),
    "ironb_scatter_rows": (_INT, [_P, _P, _I64, _INT, _P, _P]),
    "ironb_adam_step": (_INT, [_P, _INT, _I64, C.c_double, C.c_double, C.c_double, _P, _P]),
    "ironb_depth_closing": (_INT, [_P, _INT, _INT, _P, _P, _P]),
    "ironb_sobel_depth": (_INT, [_P, _INT, _INT, _P, _P]),
    "ironb_unit_dist_fwd": (_INT, [_P, _P, _P, _I64, _P, _P, _P]),
    "ironb_unit_dist_bwd": (_INT, [_P, _P, _P, _P, _P, _I64, _P, _P, _P]),
    "ironb_reparam_bwd": (_INT, [_P, _P, _P, _I64, _P, _P]),
    "ironb_matpost_fwd": (_INT, [_P, _P, _P, _I64, _INT, _P, _P, _P, _P]),
    "ironb_matpost_bwd": (_INT, [_P, _P, _P, _P, _P, _P, _I64, _INT, _P, _P, _P, _P]),
    "ironb_eik_sum_fwd": (_INT, [_P, _P, _I64, _P, _P]),
    "ironb_eik_sum_bwd": (_INT, [_P, _P, _P, _I64, _P, _P]),
    "ironb_roughrange_fwd": (_INT, [_P, _P, _I64, _F, _F, _P, _P, _P]),
    "ironb_roughrange_bwd": (_INT, [_P, _P, _P, _P, _I64, _F, _F, _P, _P]),
    "ironb_mask_rows": (_INT, [_PP, _PP, C.POINTER(C.c_int), _INT, _P, _I64, _P]),
    "ironb_neus_composite_fwd": (_INT, [_P] * 12 + [_I64, _INT, _INT, _F] + [_P] * 6 + [_P]),
    "ironb_neus_composite_bwd": (_INT, [_P] * 12 + [_I64, _INT, _INT, _F] + [_P] * 11 + [_P]),
    "ironb_neus_sections": (_INT, [_P, _P, _P, _I64, _INT, _F, _INT, _P, _P, _P, _P, _P]),
    "ironb_neus_upsample": (_INT, [_P, _P, _P, _P, _I64, _INT, _INT, _F, _P, _P, _P]),
    "ironb_neus_merge": (_INT, [_P, _P, _INT, _P, _P, _INT, _I64, _P, _P, _P]),
    "ironb_linear_fwd": (_INT, [_P, _INT, _P, _INT, _P, _INT, _INT, _INT, _INT, _P, _INT, _P]),
    "ironb_relu_mask": (_INT, [_P, _P, _I64, _P, _P]),
    "ironb_linear_wgrad": (_INT, [_P, _INT, _P, _INT, _INT, _INT, _INT, _P, _INT, _P, _P, _P]),
    "ironb_pack_tensors": (_INT, [_P, _P, _INT, _I64, _P, _F, _P]),
    "ironb_patch_loss_workspace_bytes": (_I64, [_INT, _INT, _INT]),
    "ironb_pyramid_l2": (_INT, [_P, _PI64, _P, _PI64, _INT, _INT, _INT, _P, _P, _PI64, _P, _I64, _P]),
    "ironb_ssim_loss": (_INT, [_P, _PI64, _P, _PI64, _P, _INT, _INT, _INT, _F, _INT, _F, _F, _F, _P, _P, _PI64, _P, _I64, _P]),
}

EXPORTS = tuple(_PROTOS)

_lib: Optional[C.CDLL] = None


def lib_path() -> str:
    return _build.LIB


def load() -> C.CDLL:
    """Load (building first if the in-tree .so is missing or stale).  Raises if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if _build.needs_build():
        if os.environ.get("IRONB_NO_BUILD") and os.path.exists(path):
            pass
        else:
            _build.build()
    if not os.path.exists(path):
        raise RuntimeError(f"iron_b200: {path} is missing and could not be built; there is no fallback path")
    lib = C.CDLL(path)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)   # AttributeError if the library does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


_DEBUG_SYNC = bool(os.environ.get("IRONB_DEBUG_SYNC"))


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().ironb_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"iron_b200 {what} failed (code {rc}): {msg}")
    if _DEBUG_SYNC:      # diagnostic: wait for the calling stream after every library call and name the call that faulted
        try:
            torch.cuda.current_stream().synchronize()
        except Exception as e:
            raise RuntimeError(f"iron_b200 {what}: device fault surfaced after this call: {e}") from None


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("iron_b200: tensors must live on a CUDA device (there is no CPU path)")
    if not t.is_contiguous():
        raise RuntimeError("iron_b200: internal error, non-contiguous tensor passed to a kernel")
    return t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


_WGRAD_STREAMS = {}


def wgrad_stream() -> Optional[int]:
    """A second stream for the weight-gradient products of a backward call (ironb_sdf_getall_bwd / ironb_matnet_bwd fork to
    it and join before returning): one per (device, calling stream), created on first use.  IRONB_WGRAD_STREAM=0 disables it."""
    if os.environ.get("IRONB_WGRAD_STREAM", "1") == "0":
        return None
    cur = torch.cuda.current_stream()
    key = (cur.device.index, cur.cuda_stream)
    s = _WGRAD_STREAMS.get(key)
    if s is None:
        s = torch.cuda.Stream(device=cur.device)
        _WGRAD_STREAMS[key] = s
    return s.cuda_stream


def ptr_array(ts: Sequence[Optional[torch.Tensor]]):
    arr = (C.c_void_p * len(ts))()
    for i, t in enumerate(ts):
        arr[i] = ptr(t)
    return arr


def f32c(t: torch.Tensor) -> torch.Tensor:
    """fp32 + contiguous view/copy of a tensor (no-op when already so)."""
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()
