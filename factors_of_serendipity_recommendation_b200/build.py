"""Build liblgx.so (hand-written sm_100a CUDA + the C ABI of include/lgx.h) in-tree with nvcc.

    python -m factors_of_serendipity_recommendation_b200.build [--force]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the
gpurun snapshot.
"""
from __future__ import annotations

import concurrent.futures
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OUT = os.path.join(PKG, "liblgx.so")
SOURCES = ["lgx_graph.cu", "lgx_spmm.cu", "lgx_score.cu", "lgx_score_tc.cu", "lgx_score_gq.cu", "lgx_train.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]


def _newest_source_mtime() -> float:
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(PKG, "..", "include", "lgx.h")]
    return max(os.path.getmtime(p) for p in paths)


def _compile(src: str, obj: str) -> None:
    cmd = ["nvcc", *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")


def build(force: bool = False, verbose: bool = True) -> str:
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= _newest_source_mtime():
        return OUT
    objdir = os.path.join(PKG, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = [os.path.join(objdir, s.replace(".cu", ".o")) for s in SOURCES]
    with concurrent.futures.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        list(ex.map(_compile, SOURCES, objs))
    cmd = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT, *objs,
           "-Xlinker", "--exclude-libs,ALL"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(f"built {OUT} ({os.path.getsize(OUT) / 1e6:.1f} MB)")
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv)
