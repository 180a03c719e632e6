"""Optimiser step of the training loop with mixed-precision working weights (main.py:90-94, utils.py:98-123).

The reference brackets every ``optimizer.step()`` with ``convert_models_to_fp32(model)`` / ``convert_weights(model)``:
CLIP's conv / linear / attention / projection weights live in fp16 (clip/model.py:371-392), AdamW must not see them
in that dtype (``exp_avg_sq`` underflows and ``eps`` rounds to zero: the first step yields inf/nan), so the whole
model is re-allocated in fp32, stepped, and re-allocated in fp16 -- every step.

``MasterStepper`` keeps that arithmetic (fp16 value -> fp32 -> AdamW update -> rounded back to fp16; fp32 parameters
are stepped in place) without the two whole-model re-allocations: one persistent fp32
master tensor per reduced-precision parameter, refreshed and written back with multi-tensor copies.
"""
from __future__ import annotations

from typing import Callable, Iterable, List

import torch


class MasterStepper:
    def __init__(self, params: Iterable[torch.nn.Parameter], make_optimizer: Callable[[List[torch.Tensor]], torch.optim.Optimizer]):
        self.params = [p for p in params]
        self.low = [p for p in self.params if p.dtype != torch.float32]
        self.masters = [p.detach().float().clone().requires_grad_(True) for p in self.low]
        master_of = {id(p): m for p, m in zip(self.low, self.masters)}
        # the optimiser owns fp32 tensors only: the parameter itself when it is fp32, its master otherwise
        self.opt_params = [master_of.get(id(p), p) for p in self.params]
        self.optimizer = make_optimizer(self.opt_params)

    @property
    def param_groups(self):
        return self.optimizer.param_groups

    def step(self):
        """``convert_models_to_fp32`` -> ``optimizer.step()`` -> ``convert_weights`` (main.py:90-94)."""
        live = [(p, m) for p, m in zip(self.low, self.masters) if p.grad is not None]
        if live:
            ps, ms = [p.detach() for p, _ in live], [m.detach() for _, m in live]
            torch._foreach_copy_(ms, ps)                              # fp16 value -> fp32 (utils.py:99-100)
            for p, m in live:
                if m.grad is None:
                    m.grad = torch.empty_like(m)
            torch._foreach_copy_([m.grad for _, m in live], [p.grad for p, _ in live])   # utils.py:101
        for p, m in zip(self.low, self.masters):
            if p.grad is None:
                m.grad = None
        self.optimizer.step()
        if live:
            torch._foreach_copy_(ps, ms)                              # rounded back to the working dtype (utils.py:103-123)

    def zero_grad(self, set_to_none: bool = True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()
        self.optimizer.zero_grad(set_to_none=set_to_none)
