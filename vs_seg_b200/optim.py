"""Fused multi-tensor Adam.

The reference builds ``torch.optim.Adam(model.parameters(), lr, weight_decay)``
(params/VSparams.py:388-391) and steps 178 small tensors per batch (:457-463), halving the learning
rate through ``optimizer.param_groups`` (:517-523).  ``FusedAdam`` keeps that interface (it IS a
``torch.optim.Optimizer``: ``param_groups``, ``state_dict``, ``zero_grad``, ``step``) but re-homes the
CUDA parameters of a group into ONE flat fp32 buffer (each parameter becomes a view of it, gradients
likewise) so a step is a single native launch (``vsseg_adam_step``) and a data-parallel gradient
average is a single all-reduce of the flat gradient (``vs_seg_b200.ddp``).  CPU parameters (the
reference's --debug plumbing run) use ``torch.optim.Adam`` unchanged.
"""
from __future__ import annotations

import torch

from . import lib as _lib


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("FusedAdam: invalid hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._flat = []      # per group: None (CPU group -> torch Adam) or dict(p, g, m, v, n)
        self._cpu_opt = []
        self.grad_scale = 1.0   # 1 / world_size when the flat gradients hold a SUM over ranks
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.requires_grad]
            if ps and all(p.is_cuda and p.dtype == torch.float32 for p in ps):
                _lib.load()   # no CPU fallback for CUDA parameters
                self._flat.append(self._flatten(ps))
                self._cpu_opt.append(None)
            else:
                self._flat.append(None)
                self._cpu_opt.append(torch.optim.Adam(ps, lr=group["lr"], betas=group["betas"], eps=group["eps"],
                                                      weight_decay=group["weight_decay"]))
            group["step"] = 0

    @staticmethod
    def _flatten(ps):
        dev = ps[0].device
        sizes = [(p.numel() + 3) // 4 * 4 for p in ps]   # every parameter starts 16-byte aligned
        n = sum(sizes)
        flat_p = torch.zeros(n, dtype=torch.float32, device=dev)
        flat_g = torch.zeros_like(flat_p)
        off = 0
        with torch.no_grad():
            for p, sz in zip(ps, sizes):
                view = flat_p[off:off + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view                                   # the module's parameter now lives in the flat buffer
                p.grad = flat_g[off:off + p.numel()].view(p.shape)   # autograd accumulates in place
                off += sz
        return dict(p=flat_p, g=flat_g, m=torch.zeros_like(flat_p), v=torch.zeros_like(flat_p), n=n, params=ps,
                    step_dev=torch.zeros(1, dtype=torch.int64, device=dev), lr_dev=torch.zeros(1, dtype=torch.float32, device=dev))

    # ---- torch.optim.Optimizer interface ------------------------------------------------------
    def zero_grad(self, set_to_none: bool = True):
        """The flat gradient buffers are zeroed in place (the views must survive), CPU groups follow torch."""
        for fl, opt in zip(self._flat, self._cpu_opt):
            if fl is not None:
                fl["g"].zero_()
                for p in fl["params"]:   # a user (or torch) may have dropped the view
                    if p.grad is None or p.grad.data_ptr() < fl["g"].data_ptr() or \
                            p.grad.data_ptr() >= fl["g"].data_ptr() + 4 * fl["n"]:
                        self._rebind(fl)
                        break
            else:
                opt.zero_grad(set_to_none=set_to_none)

    @staticmethod
    def _rebind(fl):
        off = 0
        for p in fl["params"]:
            g = fl["g"][off:off + p.numel()].view(p.shape)
            if p.grad is not None and p.grad.data_ptr() != g.data_ptr():
                g.copy_(p.grad)
            p.grad = g
            off += (p.numel() + 3) // 4 * 4

    def state_dict(self):
        """torch's state_dict plus the fused moments: a resumed run continues with the same exp_avg / exp_avg_sq
        (the step counters travel in param_groups), CPU groups carry their torch.optim.Adam state."""
        sd = super().state_dict()
        sd["fused"] = [None if fl is None else {"exp_avg": fl["m"].detach().clone(), "exp_avg_sq": fl["v"].detach().clone()}
                       for fl in self._flat]
        sd["cpu"] = [None if opt is None else opt.state_dict() for opt in self._cpu_opt]
        return sd

    def load_state_dict(self, state_dict):
        sd = dict(state_dict)
        fused, cpu = sd.pop("fused", None), sd.pop("cpu", None)
        super().load_state_dict(sd)
        for i, fl in enumerate(self._flat):
            if fl is not None and fused is not None and fused[i] is not None:
                if fused[i]["exp_avg"].numel() != fl["n"]:
                    raise ValueError("FusedAdam.load_state_dict: flat moment size differs from this model's parameters")
                fl["m"].copy_(fused[i]["exp_avg"])
                fl["v"].copy_(fused[i]["exp_avg_sq"])
            if fl is None and cpu is not None and cpu[i] is not None:
                self._cpu_opt[i].load_state_dict(cpu[i])

    def flat_grads(self):
        """The flat gradient buffers (one per CUDA group): what a data-parallel run all-reduces."""
        return [fl["g"] for fl in self._flat if fl is not None]

    def pre_replay(self):
        """Host side of a graph-captured step (vs_seg_b200.training.GraphedTrainStep): advance the step counters and store
        step count and learning rate in the device scalars the captured Adam launch reads."""
        for group, fl in zip(self.param_groups, self._flat):
            group["step"] += 1
            if fl is not None:
                fl["step_dev"].fill_(int(group["step"]))
                fl["lr_dev"].fill_(float(group["lr"]))

    def post_replay(self):
        """The captured launch wrote the parameters through raw pointers: bump the tensor versions (cached eval plans
        key on them)."""
        for fl in self._flat:
            if fl is not None:
                torch._C._increment_version(fl["params"])

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = None
        capturing = torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()
        for group, fl, opt in zip(self.param_groups, self._flat, self._cpu_opt):
            if not capturing:   # a captured step is counted by pre_replay() before every replay
                group["step"] += 1
            if fl is None:
                for g2 in opt.param_groups:   # lr halving is done through OUR param_groups
                    g2["lr"], g2["weight_decay"] = group["lr"], group["weight_decay"]
                opt.step()
                continue
            for p in fl["params"]:
                if p.grad is None or p.grad.data_ptr() < fl["g"].data_ptr() or \
                        p.grad.data_ptr() >= fl["g"].data_ptr() + 4 * fl["n"]:
                    self._rebind(fl)
                    break
            lib = lib or _lib.load()
            b1, b2 = group["betas"]
            s = torch.cuda.current_stream(fl["p"].device).cuda_stream
            _lib.check(lib.vsseg_adam_step(fl["p"].data_ptr(), fl["g"].data_ptr(), fl["m"].data_ptr(), fl["v"].data_ptr(),
                                           fl["n"], float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                           float(group["weight_decay"]), max(int(group["step"]), 1), float(self.grad_scale),
                                           fl["step_dev"].data_ptr() if capturing else None,
                                           fl["lr_dev"].data_ptr() if capturing else None, s),
                       "adam_step")
            _lib.count_launch()
            # the kernel wrote through raw pointers: tell torch (cached eval plans key on tensor versions)
            torch._C._increment_version(fl["params"])
        return loss
