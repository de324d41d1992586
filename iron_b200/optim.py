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
        self._rows = None            # what the pinned table currently holds
        self._table_event = None     # recorded after the last eager H2D copy of the table

    @property
    def steps_taken(self) -> int:
        return int(self._step_dev.item())

    # -- the pointer / hyper-parameter table ---------------------------------------------------------------------------
    def _table_rows(self):
        rows = []
        for p, g in self._flat:
            st = self.state[p]
            grad = p.grad
            if grad is not None and not (grad.is_contiguous() and grad.dtype == torch.float32):
                raise RuntimeError("FusedAdam: gradients must be contiguous fp32")
            rows.append((p.data_ptr(), grad.data_ptr() if grad is not None else None, st["exp_avg"].data_ptr(),
                         st["exp_avg_sq"].data_ptr(), p.numel(), float(g["lr"]), float(g["weight_decay"])))
        return rows

    def refresh_table(self) -> bool:
        """Rewrites the pinned host table from the CURRENT param_groups (lr schedules), gradients and moments; returns
        whether anything changed.  step() calls it; GraphedStage2Step.step() calls it before every replay, because the
        captured step() never runs again and the graph's copy node re-reads this pinned table at replay time.  An
        unchanged table is not rewritten; a changed one waits for the copy that may still be reading the old content."""
        rows = self._table_rows()
        if rows == self._rows:
            return False
        if self._table_event is not None and not torch.cuda.is_current_stream_capturing():
            self._table_event.synchronize()
        for t, r in zip(self._table, rows):
            t.param, t.grad, t.exp_avg, t.exp_avg_sq, t.numel, t.lr, t.weight_decay = r
        self._rows = rows
        return True

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        self.refresh_table()
        capturing = torch.cuda.is_current_stream_capturing()
        with torch.cuda.device(self._dev):
            self._table_dev.copy_(self._table_host, non_blocking=True)      # pinned -> device, stream ordered
            if not capturing:
                if self._table_event is None:
                    self._table_event = torch.cuda.Event()
                self._table_event.record()
            _lib.check(_lib.load().ironb_adam_step(_lib.ptr(self._table_dev), len(self._flat), self._max_numel,
                                                   float(self._betas[0]), float(self._betas[1]), float(self._eps),
                                                   _lib.ptr(self._step_dev), _lib.stream()), "adam_step")
        self.mark_parameters_updated()
        return loss

    def mark_parameters_updated(self) -> None:
        """The kernel writes the parameters through raw pointers, which torch's version counters do not see; everything
        keyed on `p._version` (the folded-weight cache of SDFNetwork / RenderingNetwork, autograd's saved-tensor checks)
        must learn that the values moved.  No kernel is launched.  GraphedStage2Step calls this after every replay whose
        graph contains the optimiser step."""
        bump = torch.autograd.graph.increment_version
        for p, _ in self._flat:
            bump(p)

    # -- checkpoint / resume -------------------------------------------------------------------------------------------
    def state_dict(self):
        sd = super().state_dict()
        sd["ironb_step"] = self.steps_taken        # the shared device-side step count (bias correction)
        return sd

    @torch.no_grad()
    def load_state_dict(self, state_dict):
        """Keeps the moment STORAGE (a captured graph holds pointers to it) and copies the loaded values in; restores the
        step count ("ironb_step", or the largest per-parameter `step` of a torch.optim.Adam state dict)."""
        sd = dict(state_dict)
        step = sd.pop("ironb_step", None)
        keep = {p: (self.state[p]["exp_avg"], self.state[p]["exp_avg_sq"]) for p, _ in self._flat}
        super().load_state_dict(sd)
        for p, (m, v) in keep.items():
            st = self.state[p]
            if "exp_avg" in st:
                m.copy_(st["exp_avg"])
                v.copy_(st["exp_avg_sq"])
            if step is None and "step" in st:
                step = max(int(step or 0), int(float(st["step"])))
            st["exp_avg"], st["exp_avg_sq"] = m, v
        if step is not None:
            self._step_dev.fill_(int(step))
        self._rows = None                           # param_groups were replaced: rebuild the table at the next step
