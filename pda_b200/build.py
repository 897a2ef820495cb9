"""Builds pda_b200/libpda_b200.so (sm_100a only) with nvcc; in-tree so the .so travels with gpurun.

Every .cu is compiled to its own object (in parallel, rebuilt only when it or a header changed), then linked."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
OUT = os.path.join(HERE, "libpda_b200.so")
SOURCES = ["pda_capi.cu", "pda_train.cu", "pda_step_pipe.cu", "pda_adam_lazy.cu", "pda_eval_exact.cu", "pda_eval_tc.cu",
           "pda_exchange.cu", "pda_segsum.cu", "pda_neurec.cu", "pda_debug.cu"]
HEADERS = ["pda_common.cuh", "pda_kernels.h", os.path.join("..", "..", "include", "pda_b200.h")]
CC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
            "-Xcompiler", "-O2"]
LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-Xcompiler", "-fPIC"]


def nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: pda_b200 cannot be built (there is no prebuilt or CPU fallback)")


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _headers():
    return [os.path.join(CSRC, h) for h in HEADERS]


def _obj_of(src):
    return os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")


def _obj_stale(src) -> bool:
    o = _obj_of(src)
    if not os.path.exists(o):
        return True
    t = os.path.getmtime(o)
    return any(os.path.exists(p) and os.path.getmtime(p) > t for p in [src, os.path.abspath(__file__)] + _headers())


def stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.exists(p) and os.path.getmtime(p) > t for p in sources() + _headers())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return OUT
    os.makedirs(OBJ, exist_ok=True)
    cc = nvcc()

    def compile_one(src):
        cmd = [cc] + CC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", _obj_of(src)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, r

    todo = [s for s in sources() if force or _obj_stale(s)]
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(todo)))) as ex:
        for src, r in ex.map(compile_one, todo):
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError("nvcc failed compiling " + os.path.basename(src))
            if verbose:
                print(r.stdout + r.stderr)
    r = subprocess.run([cc] + LINK_FLAGS + ["-o", OUT] + [_obj_of(s) for s in sources()], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed linking libpda_b200.so")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
