"""Helpers for the -m gpu tests: thin ctypes calls on torch tensors."""
import ctypes as C

import torch

from mfvae_b200 import _lib as L


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def gemm(engine, A, B, *, a_kmajor=True, b_kmajor=True, M=None, N=None, K=None, bias=None, epi=0, aux=None,
         c_dtype=torch.float32, C_init=None, split_k=1, groups=1):
    """C[g] (M x N) = epi(A[g] (M x K) . B[g] (N x K)^T).  A is stored [G, M, K] when a_kmajor else [G, K, M];
    B stored [G, N, K] when b_kmajor else [G, K, N]."""
    lib = L.lib()
    dt = 1 if A.dtype == torch.bfloat16 else 0
    if A.dim() == 2:
        A, B = A[None], B[None]
    G = A.shape[0]
    if a_kmajor:
        _, M_, K_ = A.shape; a_rs, a_cs = A.stride(1), 1
    else:
        _, K_, M_ = A.shape; a_rs, a_cs = 1, A.stride(1)
    if b_kmajor:
        _, N_, _ = B.shape; b_rs, b_cs = B.stride(1), 1
    else:
        _, _, N_ = B.shape; b_rs, b_cs = 1, B.stride(1)
    M, N, K = M or M_, N or N_, K or K_
    ldc = (N + 7) // 8 * 8
    Cm = torch.zeros(G, M, ldc, dtype=c_dtype, device=A.device) if C_init is None else C_init
    L.check(lib.mfvae_gemm(engine, dt, G, M, N, K, L.ptr(A), A.stride(0), a_rs, a_cs, L.ptr(B), B.stride(0), b_rs, b_cs,
                           L.ptr(Cm), Cm.stride(0), Cm.stride(1), 1 if c_dtype == torch.bfloat16 else 0,
                           L.ptr(bias), bias.stride(0) if bias is not None and bias.dim() == 2 else 0, epi,
                           L.ptr(aux), aux.stride(0) if aux is not None else 0, aux.stride(1) if aux is not None else 0,
                           split_k, stream()))
    torch.cuda.synchronize()
    return Cm[:, :, :N]


def ref_gemm(A, B, a_kmajor, b_kmajor, bias=None, epi=0, aux=None):
    if A.dim() == 2:
        A, B = A[None], B[None]
    a = A.double() if a_kmajor else A.double().transpose(1, 2)
    b = B.double() if b_kmajor else B.double().transpose(1, 2)
    c = torch.bmm(a, b.transpose(1, 2))
    if epi in (1, 2):
        c = c + (bias.double()[:, None, :] if bias.dim() == 2 else bias.double()[None, None, :])
    if epi == 2:
        c = c.clamp_min(0)
    if epi == 3:
        c = c * (aux.double() > 0)
    return c
