"""Multi-tensor Adam for the stage-2 loop.  The reference steps one torch.optim.Adam per network (six for the
'ggx' renderer: render_surface.py:112-113 and models/network_conf.py:707-716, lr 1e-5 / 1e-4 / 1e-2) -- about eight
element-wise launches per parameter tensor and step.  `FusedAdam` takes the same parameter groups and updates every tensor
of every group in ONE launch of `ironb_adam_step` with torch.optim.Adam's arithmetic (amsgrad=False).  The step counter
lives on the device, so `step()` can be captured into a CUDA graph (GraphedStage2Step(optimizer=...))."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


class _AdamTensor(C.Structure):            # mirrors ironb_adam_tensor
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("numel", C.c_int64), ("lr", C.c_float), ("weight_decay", C.c_float)]


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam(params, lr, betas, eps, weight_decay) with one kernel launch per step.

    `params`: an iterable of parameters or of param-group dicts (per-group `lr` / `weight_decay`; `betas` and `eps` are
    shared by all groups).  CUDA fp32 contiguous parameters only; amsgrad / maximize are not supported."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if not 0.0 <= lr or not 0.0 <= eps or not (0.0 <= betas[0] < 1.0 and 0.0 <= betas[1] < 1.0):
            raise ValueError("FusedAdam: invalid hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        b = {tuple(g["betas"]) for g in self.param_groups}
        e = {float(g["eps"]) for g in self.param_groups}
        if len(b) != 1 or len(e) != 1:
            raise ValueError("FusedAdam: betas and eps must be the same for every parameter group")
        self._betas, self._eps = next(iter(b)), next(iter(e))
        self._flat = [(p, g) for g in self.param_groups for p in g["params"]]
        for p, _ in self._flat:
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                raise RuntimeError("FusedAdam: parameters must be contiguous fp32 CUDA tensors (there is no CPU path)")
        self._dev = self._flat[0][0].device
        for p, _ in self._flat:
            st = self.state[p]
            st["exp_avg"] = torch.zeros_like(p)
            st["exp_avg_sq"] = torch.zeros_like(p)
        self._step_dev = torch.zeros(1, dtype=torch.int32, device=self._dev)
        n = len(self._flat)
        self._table_host = torch.empty(n * C.sizeof(_AdamTensor), dtype=torch.uint8).pin_memory()
        self._table_dev = torch.empty(n * C.sizeof(_AdamTensor), dtype=torch.uint8, device=self._dev)
        self._table = (_AdamTensor * n).from_address(self._table_host.data_ptr())
        self._max_numel = max(p.numel() for p, _ in self._flat)

    @property
    def steps_taken(self) -> int:
        return int(self._step_dev.item())

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for i, (p, g) in enumerate(self._flat):
            st = self.state[p]
            t = self._table[i]
            grad = p.grad
            if grad is not None and not (grad.is_contiguous() and grad.dtype == torch.float32):
                raise RuntimeError("FusedAdam: gradients must be contiguous fp32")
            t.param, t.grad = p.data_ptr(), (grad.data_ptr() if grad is not None else None)
            t.exp_avg, t.exp_avg_sq = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
            t.numel, t.lr, t.weight_decay = p.numel(), float(g["lr"]), float(g["weight_decay"])
        with torch.cuda.device(self._dev):
            self._table_dev.copy_(self._table_host, non_blocking=True)      # pinned -> device, stream ordered
            _lib.check(_lib.load().ironb_adam_step(_lib.ptr(self._table_dev), len(self._flat), self._max_numel,
                                                   float(self._betas[0]), float(self._betas[1]), float(self._eps),
                                                   _lib.ptr(self._step_dev), _lib.stream()), "adam_step")
        return loss
