"""In-tree build of libsoap3dp_b200.so (CUDA kernels + C ABI) for sm_100a."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsoap3dp_b200.so")
SOURCES = ["s3_index.cu", "s3_search.cu", "s3_dp.cu", "s3_seed.cu", "s3_decode.cu", "s3_params.cu", "s3_pair.cu", "s3_chain.cu", "s3_windows.cu", "s3_stages.cu", "s3_sam.cu"]
NVCC = os.environ.get("S3_NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-ccbin", "/usr/bin/g++", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wno-deprecated-declarations",
         "-Wno-deprecated-declarations", "--use_fast_math"]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "soap3dp_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, variant: str = "", defines=()) -> str:
    """variant/defines: experiment builds (kernel tuning constants as -D flags) written next to the
    product library as libsoap3dp_b200.<variant>.so; loaded only when S3_LIB_PATH names them."""
    lib = LIB if not variant else LIB.replace(".so", f".{variant}.so")
    if not variant and not force and not _stale():
        return LIB
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for s in SOURCES:
        o = os.path.join(HERE, "build", s.replace(".cu", (f".{variant}" if variant else "") + ".o"))
        cmd = [NVCC] + FLAGS + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, s), "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {s}")
    cmd = [NVCC, "-shared", "-ccbin", "/usr/bin/g++", "-Wno-deprecated-gpu-targets", "-o", lib] + objs + ["-lcudart"]
    subprocess.check_call(cmd)
    return lib


if __name__ == "__main__":
    var = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--variant=")]
    defs = [a[2:] for a in sys.argv if a.startswith("-D")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, variant=var[0] if var else "", defines=defs))
