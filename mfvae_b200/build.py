"""Build libmfvae_b200.so in-tree with nvcc for sm_100a (the only target).

    python -m mfvae_b200.build [--force]

One object per .cu (compiled in parallel), linked into ``mfvae_b200/libmfvae_b200.so``.  The .so is
git-ignored but travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libmfvae_b200.so")
SOURCES = ["comm.cu", "elementwise.cu", "enc_fused.cu", "fold.cu", "gemm_simt.cu", "gemm_tc.cu", "plan.cu", "ring.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
         "-I", os.path.join(os.path.dirname(HERE), "include")]


def _newer(dst, srcs):
    if not os.path.exists(dst):
        return False
    t = os.path.getmtime(dst)
    return all(os.path.getmtime(s) <= t for s in srcs)


def build(force=False, verbose=False):
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "mfvae.h"))
    objs, jobs = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(HERE, "build", src.replace(".cu", ".o"))
        objs.append(o)
        if force or not _newer(o, [s] + headers):
            cmd = [NVCC, *FLAGS, "-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0 or verbose:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))

    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as ex:
        list(ex.map(run, jobs))
    if jobs or force or not os.path.exists(OUT):
        run([NVCC, "-shared", "-o", OUT, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
             "-Xcompiler", "-fPIC", "--cudart", "shared", "-Xlinker", "-rpath,/usr/local/cuda/lib64", "-ldl", "-lrt", "-lpthread"])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
