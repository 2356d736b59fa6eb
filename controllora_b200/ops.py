"""Thin Python wrappers over the C ABI.  Tensors are torch CUDA tensors used purely as device-memory handles."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import GemmArgs, check

BF16 = torch.bfloat16


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _req(t: torch.Tensor, dtype, name: str):
    if not t.is_cuda:
        raise _lib.CLError(f"{name}: controllora_b200 ops need CUDA tensors (no CPU fallback)")
    if t.dtype != dtype:
        raise _lib.CLError(f"{name}: expected {dtype}, got {t.dtype}")


def gemm(
    a: torch.Tensor,
    b: torch.Tensor,
    *,
    out: Optional[torch.Tensor] = None,
    bias: Optional[torch.Tensor] = None,
    row_bias: Optional[torch.Tensor] = None,
    rows_per_group: int = 0,
    residual: Optional[torch.Tensor] = None,
    ext: Optional[torch.Tensor] = None,
    lora_up: Optional[torch.Tensor] = None,
    lora_scale: float = 1.0,
    t_add: Optional[torch.Tensor] = None,
    t_out: Optional[torch.Tensor] = None,
    out_fp32: bool = False,
    conv_stride: int = 0,
    pad_lo: int = 1,
    block_n: int = 0,
) -> torch.Tensor:
    """D = epilogue(A @ B^T).  `a` is [M, K] bf16, or an NHWC image [n, H, W, C] when conv_stride in {1, 2}
    (3x3 implicit GEMM, `b` = [N, 9*C])."""
    _req(a, BF16, "a")
    _req(b, BF16, "b")
    args = GemmArgs()
    N, K = b.shape
    if conv_stride:
        n_img, H, W, Cc = a.shape
        assert a.is_contiguous()
        args.a_mode = 1 if conv_stride == 1 else 2
        args.n_img, args.H, args.W, args.C = n_img, H, W, Cc
        args.pad_lo = pad_lo
        M = n_img * (H // conv_stride) * (W // conv_stride)
        assert K == 9 * Cc
        out_shape = (n_img, H // conv_stride, W // conv_stride, N)
    else:
        assert a.dim() == 2 and a.stride(1) == 1 and a.shape[1] == K
        M = a.shape[0]
        args.a_mode = 0
        args.lda = a.stride(0)
        out_shape = (M, N)
    assert b.stride(1) == 1
    args.M, args.N, args.K = M, N, K
    args.a = _ptr(a)
    args.b = _ptr(b)
    args.ldb = b.stride(0)
    if out is None:
        out = torch.empty(out_shape, device=a.device, dtype=torch.float32 if out_fp32 else BF16)
    else:
        _req(out, torch.float32 if out_fp32 else BF16, "out")
    out2d = out.view(-1, N) if out.is_contiguous() else out
    assert out2d.dim() == 2 and out2d.stride(1) == 1 and out2d.shape[0] == M
    args.out = _ptr(out2d)
    args.ldd = out2d.stride(0)
    args.out_fp32 = 1 if out_fp32 else 0
    if bias is not None:
        _req(bias, torch.float32, "bias")
        args.bias = _ptr(bias)
    if row_bias is not None:
        _req(row_bias, torch.float32, "row_bias")
        assert row_bias.is_contiguous() and row_bias.shape[-1] == N
        args.row_bias = _ptr(row_bias)
        args.rows_per_group = rows_per_group
    if residual is not None:
        _req(residual, BF16, "residual")
        r2 = residual.view(-1, N) if residual.is_contiguous() else residual
        assert r2.stride(1) == 1 and r2.shape[0] == M
        args.residual = _ptr(r2)
        args.ldr = r2.stride(0)
    if lora_up is not None:
        _req(lora_up, torch.float32, "lora_up")
        _req(ext, BF16, "ext")
        assert ext.shape == (16, K) and ext.stride(1) == 1
        assert lora_up.is_contiguous() and lora_up.shape[0] == N and lora_up.shape[1] in (4, 8)
        args.ext = _ptr(ext)
        args.ldb_ext = ext.stride(0)
        args.lora_up = _ptr(lora_up)
        args.lora_rp = lora_up.shape[1]
        args.lora_scale = float(lora_scale)
        if t_add is not None:
            _req(t_add, torch.float32, "t_add")
            assert t_add.is_contiguous() and t_add.shape == (M, lora_up.shape[1])
            args.t_add = _ptr(t_add)
        if t_out is not None:
            _req(t_out, torch.float32, "t_out")
            assert t_out.is_contiguous() and t_out.shape == (M, lora_up.shape[1])
            args.t_out = _ptr(t_out)
    args.block_n = block_n
    check(_lib.lib().cl_gemm(C.byref(args), _stream()), "cl_gemm")
    return out


def split_bf16_ext(down: torch.Tensor, k: int) -> torch.Tensor:
    """Pack LoRA-down rows (fp32 [r<=8, k]) into the [16, k] bf16 `ext` operand: rows j / j+8 = hi / lo split."""
    r = down.shape[0]
    assert r <= 8 and down.shape[1] == k
    ext = torch.zeros(16, k, device=down.device, dtype=BF16)
    hi = down.to(BF16)
    lo = (down - hi.float()).to(BF16)
    ext[:r] = hi
    ext[8 : 8 + r] = lo
    return ext
