"""Flat fp32 parameter / gradient / optimizer-state arenas for the trainable ControlLoRA parameters.

The reference lets DDP bucket ~400 small tensors (SURVEY.md §2.2 collective #1); here every parameter is a view into one
contiguous buffer, so data-parallel training needs exactly one all-reduce per step and the optimizer one kernel."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch


class ParamArena:
    def __init__(self, params: List[torch.nn.Parameter], device: torch.device, process_group=None):
        self.params = params
        self.pg = process_group
        self.world = 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            self.world = torch.distributed.get_world_size(process_group)
        n = sum(p.numel() for p in params)
        pad = (-n) % 4
        self.numel = n
        self.flat_p = torch.zeros(n + pad, device=device, dtype=torch.float32)
        self.flat_g = torch.zeros(n + pad, device=device, dtype=torch.float32)
        self.flat_m = torch.zeros(n + pad, device=device, dtype=torch.float32)
        self.flat_v = torch.zeros(n + pad, device=device, dtype=torch.float32)
        self._gviews: Dict[int, torch.Tensor] = {}
        off = 0
        with torch.no_grad():
            for p in params:
                k = p.numel()
                view = self.flat_p[off:off + k].view(p.shape)
                view.copy_(p.data.to(device))
                p.data = view
                self._gviews[id(p)] = self.flat_g[off:off + k].view(p.shape)
                off += k
        # the reference wraps control_lora in DDP (accelerator.prepare, train_text_to_image_control_lora.py:513), which
        # broadcasts rank 0's weights at construction: ranks that were seeded differently must not train different replicas
        if self.world > 1:
            torch.distributed.broadcast(self.flat_p, src=torch.distributed.get_global_rank(process_group, 0)
                                        if process_group is not None else 0, group=process_group)

    @property
    def grad_scale(self) -> float:
        """DDP averages gradients: the all-reduce sums, the optimizer kernel multiplies by 1/world."""
        return 1.0 / self.world

    def grad_of(self, p: torch.Tensor) -> torch.Tensor:
        return self._gviews[id(p)]

    def all_reduce(self) -> None:
        if self.world > 1:
            torch.distributed.all_reduce(self.flat_g, group=self.pg)
