"""Fused ControlLoRA training step (the body of train_text_to_image_control_lora.py:751-796 after the VAE / text
encoder): hint encoder -> UNet -> MSE -> backward -> (NCCL all-reduce) -> clip_grad_norm -> AdamW -> zero_grad, all as
kernel launches on one stream with no host synchronisation.

All trainable parameters are re-homed into ONE flat fp32 arena (and their gradients into another), so the data-parallel
gradient exchange is a single ncclAllReduce and the optimizer a single kernel.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import engine as E
from . import ops
from .engine import Ctx, Tape, Var
from .arena import ParamArena
from .hint_encoder import HintEncoderEngine

BF16 = torch.bfloat16


class Trainer:
    def __init__(self, unet, control_lora, lr: float = 1e-4, betas=(0.9, 0.999), weight_decay: float = 1e-2, eps: float = 1e-8,
                 max_grad_norm: float = 1.0, process_group=None, cuda_graph: bool = False, graph_warmup: int = 2):
        self.unet, self.cl = unet, control_lora
        self.lr, self.betas, self.wd, self.eps, self.max_norm = lr, betas, weight_decay, eps, max_grad_norm
        self.pg = process_group
        dev = unet.device_
        params: List[torch.nn.Parameter] = [p for p in control_lora.parameters() if p.requires_grad]
        seen = {id(p) for p in params}
        for p in unet.trainable_parameters():          # e.g. stacked pre_loras that are not part of control_lora
            if id(p) not in seen and p.requires_grad:
                params.append(p)
                seen.add(id(p))
        self.params = params
        self.arena = ParamArena(params, dev, process_group)
        self.world = self.arena.world
        self.flat_p, self.flat_g, self.flat_m, self.flat_v = self.arena.flat_p, self.arena.flat_g, self.arena.flat_m, self.arena.flat_v
        self.numel = self.arena.numel
        store = unet.grad_store
        store.bufs.clear()
        for p in params:
            store.bufs[id(p)] = self.arena.grad_of(p)
            store.params[id(p)] = p
        unet._runtime = None                          # rebuild the LoRA runtime against the arena views
        self.hint = HintEncoderEngine(control_lora, store.get)
        self.gnorm_sq = torch.zeros(1, device=dev, dtype=torch.float32)
        self.step_idx = 0
        self.levels = None
        # CUDA-graph mode: the forward/backward (~1800 launches) is captured once and replayed, so the step costs the
        # host one graph launch instead of ~40 ms of Python/ctypes work and cannot become launch-bound.
        self.cuda_graph = bool(cuda_graph)
        self.graph_warmup = int(graph_warmup)
        self._graph = None
        self._static = None          # static input buffers the captured kernels read
        self._static_loss = None
        self._eager_calls = 0
        self.launches_per_step = None    # kernel launches of one step (counted while capturing / running eagerly)

    def step(self, noisy_latents: torch.Tensor, timesteps: torch.Tensor, ehs: torch.Tensor, guide: torch.Tensor,
             target: torch.Tensor, eager: bool = False) -> torch.Tensor:
        """One optimizer step on device tensors: noisy_latents/target NCHW fp32, timesteps fp32 [B], ehs bf16 [B,77,768],
        guide NCHW fp32 [B,3,512,512].  Returns the (device) loss tensor; nothing is synchronised.

        With cuda_graph=True the first `graph_warmup` calls run eagerly, the next call captures the forward/backward
        into a CUDA graph (inputs are copied into static buffers first) and every later call replays it; the gradient
        all-reduce and the optimizer kernels stay outside the graph (NCCL's watchdog thread must not meet a global
        capture, and the AdamW bias correction is a host scalar).  eager=True forces the uncaptured path."""
        from . import _lib

        if not self.cuda_graph or eager:
            n0 = _lib.launch_count()
            loss = self._forward_backward(noisy_latents, timesteps, ehs, guide, target)
            self._optimizer_tail()
            self.launches_per_step = int(_lib.launch_count() - n0)
            return loss
        args = (noisy_latents, timesteps, ehs, guide, target)
        if self._static is None:
            self._static = [torch.empty_like(a_).copy_(a_) for a_ in args]
        else:
            for st, a_ in zip(self._static, args):
                if st.shape != a_.shape or st.dtype != a_.dtype:
                    raise ValueError("Trainer(cuda_graph=True): input shapes/dtypes must not change between steps")
                if st.data_ptr() != a_.data_ptr():
                    st.copy_(a_, non_blocking=True)
        if self._graph is None and self._eager_calls < self.graph_warmup:
            self._eager_calls += 1
            loss = self._forward_backward(*self._static)
            self._optimizer_tail()
            return loss
        if self._graph is None:
            graph = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            try:
                with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                    self._static_loss = self._forward_backward(*self._static)
            except Exception as ex:      # capture is an optimisation of the launch path only: say so and keep training eagerly
                import sys

                print(f"controllora_b200.Trainer: CUDA-graph capture failed ({type(ex).__name__}: {ex}); "
                      f"continuing with per-kernel launches", file=sys.stderr, flush=True)
                torch.cuda.synchronize()
                self.cuda_graph = False
                self._static_loss = None
                loss = self._forward_backward(*self._static)
                self._optimizer_tail()
                return loss
            self.launches_per_step = int(_lib.launch_count() - n0) + 2   # + sumsq + adamw outside the graph
            self._graph = graph
        self._graph.replay()
        self._optimizer_tail()
        return self._static_loss

    def _forward_backward(self, noisy_latents, timesteps, ehs, guide, target) -> torch.Tensor:
        tape = Tape()
        hctx = Ctx(tape=tape)
        states = self.hint.forward(hctx, guide)
        control = {}
        for procs, s in zip(self.cl.lora_layers, states):
            n, H, W, C = s.data.shape
            c = Var(s.data.view(n, H * W, C), rg=True)       # token-matrix view of the same memory for the UNet side
            for proc in procs:
                proc.control_states = c.data
            control[c.data.data_ptr()] = c

            def bridge(s=s, c=c, shape=(n, H, W, C)):
                if c.grad is not None:
                    E.give_tensor(s, c.grad.view(shape))
                    c.grad = None

            tape.record(bridge)      # runs after every UNet backward op (incl. the per-level d-control GEMMs)
        pred, ctx, rt = self.unet.run_engine(noisy_latents, timesteps, ehs, control, tape)
        loss, dpred = ops.mse_loss(pred.data, target)
        pred.grad = dpred
        tape.backward()
        return loss

    def _optimizer_tail(self) -> None:
        self.step_idx += 1
        self.arena.all_reduce()          # one ncclAllReduce(sum) over the whole gradient arena
        self.gnorm_sq.zero_()
        ops.sumsq(self.flat_g, self.gnorm_sq)
        ops.adamw(self.flat_p, self.flat_g, self.flat_m, self.flat_v, self.lr, self.betas[0], self.betas[1], self.eps, self.wd,
                  self.step_idx, gnorm_sq=self.gnorm_sq, max_norm=self.max_norm, grad_scale=self.arena.grad_scale, zero_grad=True)

    # ------------------------------------------------------------------------------------------------ checkpoints
    # The reference checkpoints through accelerate: `accelerator.save_state(output_dir/checkpoint-{global_step})` every
    # `checkpointing_steps` (train_text_to_image_control_lora.py:805-809) and resumes from the highest `checkpoint-N`
    # (:713-735).  Same directory convention here: the ControlLoRA weights in the reference's own format
    # (config.json + diffusion_pytorch_model.safetensors, the files of :927-929) plus the optimizer state of the flat arenas.
    def save_checkpoint(self, output_dir, global_step: Optional[int] = None) -> str:
        import os

        step = self.step_idx if global_step is None else int(global_step)
        path = os.path.join(str(output_dir), f"checkpoint-{step}")
        os.makedirs(path, exist_ok=True)
        self.cl.save_config(path)
        self.cl.save_pretrained(path, safe_serialization=True)
        torch.save({"step_idx": self.step_idx, "global_step": step, "numel": self.numel, "lr": self.lr, "betas": self.betas,
                    "weight_decay": self.wd, "eps": self.eps, "max_grad_norm": self.max_norm,
                    "exp_avg": self.flat_m[:self.numel].detach().cpu(), "exp_avg_sq": self.flat_v[:self.numel].detach().cpu(),
                    "param_names": [n for n, _ in self.cl.named_parameters()]},
                   os.path.join(path, "optimizer.bin"))
        return path

    @staticmethod
    def latest_checkpoint(output_dir) -> Optional[str]:
        """Highest `checkpoint-N` under output_dir (train_text_to_image_control_lora.py:713-721), or None."""
        import os

        if not os.path.isdir(str(output_dir)):
            return None
        best, best_n = None, -1
        for d in os.listdir(str(output_dir)):
            if d.startswith("checkpoint-") and d[len("checkpoint-"):].isdigit() and int(d[len("checkpoint-"):]) > best_n:
                best, best_n = d, int(d[len("checkpoint-"):])
        return None if best is None else os.path.join(str(output_dir), best)

    def load_checkpoint(self, path) -> int:
        """Restore parameters (into the flat arena the kernels read) and AdamW moments; returns the stored global step.
        A captured CUDA graph stays valid: it reads the same arena addresses."""
        import os

        st = os.path.join(str(path), "diffusion_pytorch_model.safetensors")
        if os.path.isfile(st):
            from safetensors.torch import load_file

            sd = load_file(st)
        else:
            sd = torch.load(os.path.join(str(path), "diffusion_pytorch_model.bin"), map_location="cpu")
        own = dict(self.cl.named_parameters())
        missing = [k for k in own if k not in sd]
        if missing:
            raise KeyError(f"checkpoint {path} lacks parameters: {missing[:4]}...")
        with torch.no_grad():
            for k, p in own.items():
                p.data.copy_(sd[k].to(p.data.device, p.data.dtype))      # p.data is a view into flat_p
        opt = torch.load(os.path.join(str(path), "optimizer.bin"), map_location="cpu")
        if int(opt["numel"]) != self.numel:
            raise ValueError("optimizer state does not match this ControlLoRA's parameter count")
        with torch.no_grad():
            self.flat_m[:self.numel].copy_(opt["exp_avg"].to(self.flat_m.device))
            self.flat_v[:self.numel].copy_(opt["exp_avg_sq"].to(self.flat_v.device))
            self.flat_g.zero_()
        self.step_idx = int(opt["step_idx"])
        return int(opt.get("global_step", self.step_idx))
