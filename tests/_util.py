"""Shared helpers of the GPU parity tests (call liboasr through its C-ABI via ctypes)."""
import ctypes as C

import numpy as np
import torch

from omnilingual_asr import _native as N


def lib():
    return N.load()


def bf16_round(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


def sync():
    torch.cuda.synchronize()


def gemm(A, W, bias, epilogue, out=None, resid=None, ln_g=None, ln_b=None, keys=None, ldo=None):
    M, K = A.shape
    Nn = W.shape[0]
    rc = lib().oasr_gemm(N.ptr(A), N.ptr(W), N.ptr(bias), M, Nn, K, epilogue, N.ptr(out),
                         ldo if ldo is not None else (out.shape[1] if out is not None else Nn), N.ptr(resid),
                         N.ptr(ln_g), N.ptr(ln_b), N.ptr(keys), N.stream_ptr())
    N.check(rc, "oasr_gemm")
    sync()
    return out


def unpack_keys(keys: torch.Tensor) -> np.ndarray:
    k = keys.cpu().numpy().astype(np.uint64)
    return (np.uint64(0xFFFFFFFF) - (k & np.uint64(0xFFFFFFFF))).astype(np.int64)


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.float().cpu()
    b = b.float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
