"""Fused Adam over a flat parameter arena (hulc2.py:185-198, conf/model/optimizer/adam.yaml).

``torch.optim.Adam`` semantics (amsgrad=False, maximize=False), but parameters, gradients and both
moments live in four contiguous fp32 arenas so one kernel launch updates all 47 M parameters and a
data-parallel reducer can all-reduce contiguous gradient buckets.  ``param.data`` / ``param.grad``
become views into the arenas; state_dict names/shapes of the module are unchanged.
"""
from __future__ import annotations

from typing import Dict, Iterable, List

import torch

from ._lib import call
from ._lib import tag as _lib_tag

_ALIGN = 8  # elements: every parameter starts 32-byte aligned, so its bf16 mirror (ops.weight16) is a legal TMA base


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params: Iterable, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self._arenas: List[Dict] = []
        self.grad_scale = 1.0  # multiplied into gradients inside the kernel (e.g. 1/world_size)
        self.capturable = False  # True: step number and learning rate are read from device memory (CUDA-graph replays)
        self._step_dev = None
        self._lr_dev = None      # float32[n_groups] on the device, mirrors group["lr"] (see sync_lr)
        self._lr_host: List[float] = []
        for group in self.param_groups:
            self._arenas.append(self._build_arena(group))

    # ------------------------------------------------------------------ arena
    @staticmethod
    def _build_arena(group) -> Dict:
        ps = [p for p in group["params"] if p.requires_grad]
        if not ps:
            return {"params": [], "n": 0}
        dev = ps[0].device
        offs, n = [], 0
        for p in ps:
            if p.dtype != torch.float32:
                raise TypeError("FusedAdam keeps fp32 master parameters")
            offs.append(n)
            n += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        arena = {"params": ps, "offs": offs, "n": n}
        for name in ("p", "g", "m", "v"):
            arena[name] = torch.zeros(n, device=dev, dtype=torch.float32)
        with torch.no_grad():
            for p, o in zip(ps, offs):
                view = arena["p"][o : o + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
        arena["step"] = 0
        arena["views"] = {id(p): arena["g"][o : o + p.numel()].view(p.shape) for p, o in zip(ps, offs)}
        from . import ops

        ops.register_param_arena(arena["p"])   # one flat f32->bf16 conversion per step serves every contraction
        # gradient sinks: backward kernels write the first gradient of a parameter straight into its arena slice
        import weakref

        keys = ops.register_grad_views(arena["g"], ps, offs)
        arena["_fin"] = weakref.finalize(arena["g"], ops.unregister_grad_views, keys)
        return arena

    def _attach_grads(self, arena) -> None:
        """Make every ``param.grad`` the matching view of the gradient arena (copying stray grads in)."""
        for p, o in zip(arena["params"], arena["offs"]):
            view = arena["g"][o : o + p.numel()].view(p.shape)
            if p.grad is None:
                p.grad = view
            elif p.grad.data_ptr() != view.data_ptr():
                g = p.grad.contiguous()
                call("hulc2_copy2d", g.data_ptr(), g.numel(), view.data_ptr(), g.numel(), 1, g.numel(), 0)
                p.grad = view

    def attach(self) -> None:
        for a in self._arenas:
            if a["n"]:
                self._attach_grads(a)

    def land_grad(self, p) -> None:
        """Make ``p.grad`` the arena slice of ``p`` (copying a gradient that autograd adopted from elsewhere).  Called from
        the reducer's post-accumulate hook so that every gradient is in the arena before its bucket is all-reduced."""
        for a in self._arenas:
            view = a["views"].get(id(p)) if a["n"] else None
            if view is not None:
                if p.grad is not None and p.grad.data_ptr() != view.data_ptr():
                    g = p.grad.contiguous()
                    if g.is_cuda:
                        call("hulc2_copy2d", g.data_ptr(), g.numel(), view.data_ptr(), g.numel(), 1, g.numel(), 0)
                    else:                       # host tensors only occur in the gloo reducer tests
                        view.copy_(g)
                    p.grad = view
                return

    def grad_arenas(self) -> List[torch.Tensor]:
        """Gradient arenas.  NOTE for data-parallel runs: after ``GradBucketReducer.finish()`` they (and every ``param.grad``)
        hold the SUM over ranks; the 1/world mean is applied inside the Adam kernel (``grad_scale``).  Scale by
        ``grad_scale`` before clipping or logging gradient norms."""
        return [a["g"] for a in self._arenas if a["n"]]

    # ------------------------------------------------------------------ optimizer API
    def zero_grad(self, set_to_none: bool = False) -> None:  # noqa: D401
        """Zero-fills the gradient arena.  ``param.grad`` is reset to None: the backward kernels write the first gradient
        of every parameter directly into its arena slice (``ops.grad_buffer``) and autograd adopts that tensor as
        ``param.grad`` without an accumulate kernel; ``step()`` / ``attach()`` re-attach whatever did not arrive that way
        (parameters without a gradient this step read as zeros)."""
        from . import ops

        for a in self._arenas:
            if a["n"]:
                if a["g"].is_cuda:
                    call("hulc2_fill", a["g"].data_ptr(), a["n"], 0.0)
                else:                           # host tensors only occur in the gloo reducer tests
                    a["g"].zero_()
                for p in a["params"]:
                    p.grad = None
        ops.begin_grad_step()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self.capturable:
            if not torch.cuda.is_current_stream_capturing():
                self.sync_lr()              # (a fill captured into a graph would pin the learning rate at its capture-time value)
            elif self._lr_dev is None or any(h != float(g["lr"]) for h, g in zip(self._lr_host, self.param_groups)):
                raise RuntimeError("FusedAdam: call sync_lr() before capturing a step")
        for gi, (group, a) in enumerate(zip(self.param_groups, self._arenas)):
            if not a["n"]:
                continue
            self._check_bound(a)
            self._attach_grads(a)
            ctr = self.step_counter(a["p"].device) if self.capturable else None
            a["step"] += 1
            b1, b2 = group["betas"]
            if self.capturable:
                # step = (device counter of completed steps) + 1; the trainer bumps the counter after every step.  The
                # learning rate is read from device memory, so schedulers keep working when this launch is replayed.
                _lib_tag(f"adam_step[n={a['n']}]", 0.0, 28.0 * a["n"])      # 16 B/param read + 12 B/param written
                call("hulc2_adam_step_graph", a["p"].data_ptr(), a["g"].data_ptr(), a["m"].data_ptr(), a["v"].data_ptr(), a["n"],
                     self._lr_dev.data_ptr() + 4 * gi, float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]),
                     ctr.data_ptr(), 1, float(self.grad_scale))
            else:
                _lib_tag(f"adam_step[n={a['n']}]", 0.0, 28.0 * a["n"])
                call("hulc2_adam_step", a["p"].data_ptr(), a["g"].data_ptr(), a["m"].data_ptr(), a["v"].data_ptr(), a["n"],
                     float(group["lr"]), float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]), int(a["step"]),
                     float(self.grad_scale))
        from . import ops

        ops.invalidate_weight_mirrors()
        return loss

    @staticmethod
    def _check_bound(a) -> None:
        """``param.data`` must still be the arena view: ``model.to()/.half()`` after construction rebinds it and the update
        would silently go nowhere."""
        lo = a["p"].data_ptr()
        hi = lo + 4 * a["n"]
        for p in a["params"]:
            if not (lo <= p.data_ptr() < hi) or p.dtype != torch.float32:
                raise RuntimeError("FusedAdam: a parameter no longer lives in the optimizer's arena (model.to()/.half() after "
                                   "configure_optimizers?) -- build the optimizer after moving the model")

    def captured_scalars(self) -> tuple:
        """Host scalars a captured step freezes into kernel arguments (everything but lr and the step number)."""
        return tuple((tuple(g["betas"]), float(g["eps"]), float(g["weight_decay"])) for g in self.param_groups) + (float(self.grad_scale),)

    def sync_lr(self) -> None:
        """Mirror ``group["lr"]`` into the device vector the capturable Adam launch reads (one 4-byte fill per changed group,
        on the current stream -- ordered before the step / replay that follows)."""
        dev = next((a["p"].device for a in self._arenas if a["n"]), None)
        if dev is None:
            return
        if self._lr_dev is None:
            self._lr_dev = torch.zeros(len(self.param_groups), dtype=torch.float32, device=dev)
            self._lr_host = [None] * len(self.param_groups)
        for gi, group in enumerate(self.param_groups):
            lr = float(group["lr"])
            if self._lr_host[gi] != lr:
                call("hulc2_fill", self._lr_dev.data_ptr() + 4 * gi, 1, lr)
                self._lr_host[gi] = lr

    def step_counter(self, device) -> torch.Tensor:
        """int64[1] device counter of completed optimizer steps (used when ``capturable``)."""
        if self._step_dev is None:
            done = max((a["step"] for a in self._arenas if a["n"]), default=0)
            self._step_dev = torch.full((1,), int(done), dtype=torch.int64, device=device)
        return self._step_dev

    def note_replayed_step(self) -> None:
        """A captured step was replayed: keep the host-side step numbers (state_dict) in sync with the device counter."""
        for a in self._arenas:
            if a["n"]:
                a["step"] += 1

    # ------------------------------------------------------------------ checkpointing (torch.optim.Adam-compatible layout)
    def state_dict(self):
        state, idx = {}, 0
        groups = []
        for group, a in zip(self.param_groups, self._arenas):
            ids = []
            for p in group["params"]:
                ids.append(idx)
                if a["n"] and any(p is q for q in a["params"]):
                    o = a["offs"][[i for i, q in enumerate(a["params"]) if q is p][0]]
                    state[idx] = {
                        "step": torch.tensor(float(a["step"])),
                        "exp_avg": a["m"][o : o + p.numel()].view(p.shape).clone(),
                        "exp_avg_sq": a["v"][o : o + p.numel()].view(p.shape).clone(),
                    }
                idx += 1
            g = {k: v for k, v in group.items() if k != "params"}
            g["params"] = ids
            groups.append(g)
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd):
        idx = 0
        for group, a, g_sd in zip(self.param_groups, self._arenas, sd["param_groups"]):
            for k, v in g_sd.items():
                if k != "params":
                    group[k] = v
            for p in group["params"]:
                st = sd["state"].get(idx, sd["state"].get(str(idx)))
                if st is not None and a["n"]:
                    o = a["offs"][[i for i, q in enumerate(a["params"]) if q is p][0]]
                    a["m"][o : o + p.numel()].view(p.shape).copy_(st["exp_avg"])
                    a["v"][o : o + p.numel()].view(p.shape).copy_(st["exp_avg_sq"])
                    a["step"] = int(float(st["step"]))
                idx += 1
        if self._step_dev is not None:          # resume: the device counter of completed steps follows the loaded state
            done = max((a["step"] for a in self._arenas if a["n"]), default=0)
            self._step_dev.fill_(int(done))
        self._lr_host = [None] * len(self._lr_host)
