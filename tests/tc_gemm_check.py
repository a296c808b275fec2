"""Standalone checker of the tcgen05 GEMM against an fp64 torch reference (run in its own process so that a
device-side trap cannot poison the caller's CUDA context).  Prints one line per case and exits non-zero on the
first mismatch.    python tests/tc_gemm_check.py [quick|full]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.gpu_util import gemm, ref_gemm   # noqa: E402

TC = 2


def case(name, G, M, N, K, akm, bkm, epi=0, c_dtype=torch.float32, split_k=1, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    def mk(rows, cols, kmajor):
        shape = (G, rows, cols) if kmajor else (G, cols, rows)
        ld = (shape[2] + 7) // 8 * 8
        t = torch.randn(shape[0], shape[1], ld, device="cuda", generator=g).to(torch.bfloat16)
        return t[:, :, :shape[2]]
    A, B = mk(M, K, akm), mk(N, K, bkm)
    bias = torch.randn(G, (N + 7) // 8 * 8, device="cuda", generator=g) if epi in (1, 2) else None
    aux = None
    if epi == 3:
        aux = torch.randn(G, M, (N + 7) // 8 * 8, device="cuda", generator=g).to(torch.bfloat16)
    C0 = None
    if epi == 4:
        C0 = torch.zeros(G, M, (N + 7) // 8 * 8, device="cuda")
    got = gemm(TC, A, B, a_kmajor=akm, b_kmajor=bkm, M=M, N=N, K=K, bias=bias, epi=epi, aux=aux, c_dtype=c_dtype,
               C_init=C0, split_k=split_k)
    want = ref_gemm(A, B, akm, bkm, bias[:, :N] if bias is not None else None, epi, aux[:, :, :N] if aux is not None else None)
    if epi == 4:
        want = ref_gemm(A, B, akm, bkm)
    err = (got.double() - want).abs().max().item()
    scale = want.abs().max().item() + 1e-9
    tol = 1.2e-2 if c_dtype == torch.bfloat16 else 2e-3
    ok = err <= tol * scale
    print(f"{'ok  ' if ok else 'FAIL'} {name}: G={G} M={M} N={N} K={K} A={'K' if akm else 'MN'} B={'K' if bkm else 'MN'} "
          f"epi={epi} err={err:.3e} scale={scale:.3e}", flush=True)
    return ok


def main(mode):
    ok = True
    # forward: K-major x K-major
    ok &= case("fwd-basic", 1, 128, 64, 64, True, True)
    ok &= case("fwd-k256", 1, 256, 128, 256, True, True, epi=1)
    ok &= case("fwd-bias-relu-bf16", 1, 512, 256, 1024, True, True, epi=2, c_dtype=torch.bfloat16)
    ok &= case("fwd-ragged", 1, 200, 72, 208, True, True, epi=1)
    ok &= case("fwd-grouped", 5, 384, 64, 208, True, True, epi=2, c_dtype=torch.bfloat16)
    ok &= case("fwd-N40-K40", 1, 256, 40, 40, True, True, epi=1)
    # dgrad: K-major x MN-major
    ok &= case("dgrad-basic", 1, 128, 64, 64, True, False)
    ok &= case("dgrad-mask", 2, 384, 256, 64, True, False, epi=3, c_dtype=torch.bfloat16)
    ok &= case("dgrad-ragged", 1, 256, 5120 // 8, 72, True, False, c_dtype=torch.bfloat16)
    # wgrad: MN-major x MN-major, fp32 accumulate, split-K
    ok &= case("wgrad-basic", 1, 128, 64, 64, False, False, epi=4)
    ok &= case("wgrad-splitk", 1, 256, 208, 2048, False, False, epi=4, split_k=4)
    ok &= case("wgrad-small", 3, 64, 64, 1024, False, False, epi=4)
    ok &= case("wgrad-ragged", 1, 40, 1024, 512, False, False, epi=4)
    if mode == "full":
        ok &= case("fwd-big", 1, 4096, 2048, 5120, True, True, epi=2, c_dtype=torch.bfloat16)
        ok &= case("fwd-5660", 1, 4096, 5660, 1024, True, True, epi=1)
        ok &= case("dgrad-big", 1, 4096, 5120, 2048, True, False, c_dtype=torch.bfloat16)
        ok &= case("dgrad-5660", 1, 4096, 1024, 5660, True, False, epi=3, c_dtype=torch.bfloat16)
        ok &= case("wgrad-big", 1, 2048, 5120, 4096, False, False, epi=4)
        ok &= case("wgrad-5660", 1, 5660, 1024, 4096, False, False, epi=4)
        ok &= case("enc-fwd", 40, 4096, 256, 64, True, True, epi=2, c_dtype=torch.bfloat16)
        ok &= case("enc-wgrad", 40, 128, 256, 4096, False, False, epi=4)
    print("TC_GEMM_ALL_OK" if ok else "TC_GEMM_FAILED", flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main(sys.argv[1] if len(sys.argv) > 1 else "quick"))
