"""ctypes wrappers of the fused policy-net kernels in libmjx (csrc/mjx_nn.cuh). Tensors are CUDA bf16, logically
[B, C, 1, L] in channels_last memory format, i.e. [B, L, C] in memory."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def _stream(t: torch.Tensor):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _check_nhwc(x: torch.Tensor):
    assert x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4 and x.shape[2] == 1
    b, c, _, l = x.shape
    assert c % 8 == 0 and x.stride(1) == 1 and x.stride(3) == c and x.stride(0) == c * l, "expected channels_last [B, C, 1, L]"
    return b, c, l


def affine_mish(x: torch.Tensor, scale: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """mish(x * scale[c] + bias[c]); scale / bias float32 [C]"""
    b, c, l = _check_nhwc(x)
    assert scale.dtype == torch.float32 and bias.dtype == torch.float32 and scale.numel() == c == bias.numel()
    out = torch.empty_like(x)
    _lib.check(_lib.load().mjx_nn_affine_mish_bf16(C.c_void_p(x.data_ptr()), C.c_void_p(scale.data_ptr()), C.c_void_p(bias.data_ptr()),
                                                   C.c_void_p(out.data_ptr()), x.numel(), c, _stream(x)), "mjx_nn_affine_mish_bf16")
    return out


def pool_mean_max(x: torch.Tensor):
    """(mean over L, max over L) -> two bf16 [B, C] tensors"""
    b, c, l = _check_nhwc(x)
    avg = torch.empty((b, c), dtype=torch.bfloat16, device=x.device)
    mx = torch.empty((b, c), dtype=torch.bfloat16, device=x.device)
    _lib.check(_lib.load().mjx_nn_pool_bf16(C.c_void_p(x.data_ptr()), C.c_void_p(avg.data_ptr()), C.c_void_p(mx.data_ptr()), b, l, c,
                                            _stream(x)), "mjx_nn_pool_bf16")
    return avg, mx


def gate_residual(y: torch.Tensor, gate: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """y * gate[b, c] + x; gate bf16 [B, C] contiguous"""
    b, c, l = _check_nhwc(y)
    assert _check_nhwc(x) == (b, c, l) and gate.dtype == torch.bfloat16 and gate.shape == (b, c) and gate.is_contiguous()
    out = torch.empty_like(y)
    _lib.check(_lib.load().mjx_nn_gate_residual_bf16(C.c_void_p(y.data_ptr()), C.c_void_p(gate.data_ptr()), C.c_void_p(x.data_ptr()),
                                                     C.c_void_p(out.data_ptr()), b, l, c, _stream(y)), "mjx_nn_gate_residual_bf16")
    return out
