"""Builds ssr-speech_b200/libssr_b200.so (sm_100a only) with nvcc, in-tree."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libssr_b200.so")
SOURCES = ["gemm_simt.cu", "gemm_tc.cu", "gemm_layer.cu", "gemm_flat2.cu", "lm_kernels.cu", "lm_mega.cu", "attn_decode_tma.cu", "attn_prefill_mma.cu", "lm_engine.cu", "codec_kernels.cu", "codec_cl.cu", "conv_tc.cu", "conv_tc32.cu", "resblock_tc.cu", "codec_engine.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--use_fast_math=false",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isfile(c) or c == "nvcc"):
            return c
    return "nvcc"


def needs_build() -> bool:
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "ssr_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.isfile(d))


def build(force: bool = False, verbose: bool = False, extra_flags=()) -> str:
    if not force and not needs_build():
        return LIB
    objs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    procs = []
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        if not os.path.isfile(sp):
            continue
        obj = os.path.join(HERE, "build", src.replace(".cu", ".o"))
        objs.append(obj)
        if not force and os.path.isfile(obj) and os.path.getmtime(obj) > max(
                os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC)) and \
                os.path.getmtime(obj) > os.path.getmtime(os.path.join(HERE, "..", "include", "ssr_b200.h")):
            continue
        cmd = [_nvcc(), *[f for f in NVCC_FLAGS if f != "--use_fast_math=false"], *extra_flags, "-c", sp, "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{out}")
        if verbose and out.strip():
            print(out)
    cmd = [_nvcc(), "-shared", "-o", LIB, *objs, "-lcudart"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    return LIB


def build_variant(out: str, defines) -> str:
    """A second library with extra -D macros (e.g. LK_STRONG_SYNC=1, DEC_STAGES=3) for same-box A/B runs: select it with
    SSRB_LIB=<out> (tools/gpu_ab.sh).  Objects go to a scratch directory; the in-tree library is untouched."""
    import tempfile
    tmp = tempfile.mkdtemp(prefix="ssrb_variant_")
    objs, procs = [], []
    for src in SOURCES:
        obj = os.path.join(tmp, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [_nvcc(), *[f for f in NVCC_FLAGS if f != "--use_fast_math=false"], *[f"-D{d}" for d in defines], "-c",
               os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        o, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{o}")
    r = subprocess.run([_nvcc(), "-shared", "-o", out, *objs, "-lcudart"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    return out


if __name__ == "__main__":
    if "--variant" in sys.argv:          # python build.py --variant out.so LK_STRONG_SYNC=1 [MORE=...]
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
    else:
        print(build(force="--force" in sys.argv, verbose=True, extra_flags=["-Xptxas", "-v"] if "--ptxas" in sys.argv else []))
