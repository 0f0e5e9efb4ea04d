"""Fused per-step tail of the reference's training loop (engine/training_engine.py:289-312) for a module whose parameters have been
flattened into one fp32 buffer: GradScaler unscale + inf check, ``clip_grad_norm_``, AdamW with the reference's two parameter groups
(cvnets/misc/common.py:122-176), GradScaler update -- two kernel launches, all state on the device (so the step stays one CUDA graph).

    tail = FlatAdamW(model, lr=2e-3, weight_decay=0.05, max_norm=10.0)
    loss = criterion(model(x), y); model.zero_grad(set_to_none=True)
    (loss * tail.loss_scale()).backward()
    tail.step()            # gathers .grad into the flat buffer (optionally all-reduces it), then cvb_grad_norm + cvb_adamw_step
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib as L
from .ops import _count, _lib, _stream


class FlatAdamW:
    def __init__(self, model: torch.nn.Module, lr: float = 2e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.05,
                 no_decay_bn_filter_bias: bool = True, max_norm: float = 10.0, init_scale: float = 65536.0, growth_factor: float = 2.0,
                 backoff_factor: float = 0.5, growth_interval: int = 2000):
        self.params = [p for p in model.parameters() if p.requires_grad]
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.n = n
        self.flat_p = torch.empty(n, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(n, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros(n, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(n, device=dev, dtype=torch.float32)
        self.wd = torch.empty(n, device=dev, dtype=torch.float32)
        o = 0
        for p in self.params:
            k = p.numel()
            assert p.dtype == torch.float32 and p.is_contiguous()
            self.flat_p[o:o + k].copy_(p.data.view(-1))
            p.data = self.flat_p[o:o + k].view_as(p)  # parameters become views of the flat buffer (state_dict / modules see no change)
            decay = weight_decay if not (no_decay_bn_filter_bias and p.dim() == 1) else 0.0
            self.wd[o:o + k].fill_(decay)
            o += k
        self.stats = torch.zeros(4, device=dev, dtype=torch.float32)
        self.scale = torch.tensor([init_scale, 0.0], device=dev, dtype=torch.float32)
        self.step_count = torch.zeros(1, device=dev, dtype=torch.float32)
        self.hp = (float(lr), float(betas[0]), float(betas[1]), float(eps), float(max_norm))
        self.gs = (float(growth_factor), float(backoff_factor), int(growth_interval))

    def loss_scale(self) -> torch.Tensor:
        return self.scale[0]

    def gather_grads(self) -> torch.Tensor:
        """.grad tensors -> the flat gradient buffer (one batched copy); parameters without a gradient contribute zeros."""
        views = []
        for p in self.params:
            views.append(p.grad.reshape(-1) if p.grad is not None else torch.zeros(p.numel(), device=p.device, dtype=torch.float32))
        torch.cat(views, out=self.flat_g)
        return self.flat_g

    def step(self, world: int = 1, all_reduce=None) -> None:
        g = self.gather_grads()
        if world > 1 and all_reduce is not None:
            all_reduce(g)  # SUM over ranks; the mean is folded into the unscale below by the caller's choice of all_reduce
        lib = _lib()
        L.check(lib.cvb_grad_norm(g.data_ptr(), self.n, self.scale.data_ptr(), self.stats.data_ptr(), _stream()), "cvb_grad_norm")
        _count()
        lr, b1, b2, eps, max_norm = self.hp
        gf, bf, gi = self.gs
        L.check(lib.cvb_adamw_step(self.flat_p.data_ptr(), g.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), self.wd.data_ptr(),
                                   self.n, lr, b1, b2, eps, max_norm, self.stats.data_ptr(), self.scale.data_ptr(), self.step_count.data_ptr(),
                                   gf, bf, gi, _stream()), "cvb_adamw_step")
        _count()
