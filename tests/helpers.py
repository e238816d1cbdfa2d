"""Shared test helpers (oracle loading, error metrics)."""
import ctypes as C
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a-b|| / ||b|| in fp64 (b = oracle)."""
    a = a.detach().double().cpu().flatten()
    b = b.detach().double().cpu().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def load_vq_oracle():
    path = os.path.join(ROOT, "oracle", "_build", "libvq_argmin_ref.so")
    if not os.path.exists(path):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    lib = C.CDLL(path)
    for fn in (lib.vq_argmin_ref, lib.vq_argmin_ref_order1):
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        fn.restype = None
    return lib


def vq_oracle(z: np.ndarray, e: np.ndarray, order: int = 0):
    """Returns (idx int64 [N], best [N], second [N]) from oracle/vq_argmin_ref.c (order 0: one FMA chain over d;
    order 1: even/odd chains, the summation order of the packed-FMA kernel)."""
    lib = load_vq_oracle()
    z = np.ascontiguousarray(z, dtype=np.float32)
    e = np.ascontiguousarray(e, dtype=np.float32)
    N, D = z.shape
    K = e.shape[0]
    idx = np.zeros(N, dtype=np.int64)
    best = np.zeros(N, dtype=np.float32)
    second = np.zeros(N, dtype=np.float32)
    fn = lib.vq_argmin_ref if order == 0 else lib.vq_argmin_ref_order1
    fn(z.ctypes.data, e.ctypes.data, idx.ctypes.data, best.ctypes.data, second.ctypes.data, N, K, D)
    return idx, best, second


def tf32_round(x: torch.Tensor) -> torch.Tensor:
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def tf32_trunc(x: torch.Tensor) -> torch.Tensor:
    return (x.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)
