"""Build libsfe.so with nvcc for sm_100a.  In-tree, no JIT cache.

The library is one object for the C ABI / small kernels (sfe_lib.cu) plus one object per FFT plan
(sfe_plans.cu compiled with -DSFE_PLAN_TU=i), compiled in parallel and linked into one shared library.
SFE_DEV_PLANS="12,13" (development only) recompiles just those plan objects when only kernel headers changed.
"""
import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
HEADERS = sorted(f for f in os.listdir(HERE) if f.endswith((".cuh", ".h"))) + [os.path.join("..", "..", "include", "sfe.h")]
SOURCES = ["sfe_lib.cu", "sfe_plans.cu"]
OUT = os.path.join(HERE, "libsfe.so")
OBJ = os.path.join(HERE, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def num_plans() -> int:
    txt = open(os.path.join(HERE, "sfe_registry.h")).read()
    return len(re.findall(r"^\s*X\(\d+,", txt, flags=re.M))


def stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(HERE, f)) > t for f in SOURCES + HEADERS)


def _nvcc() -> str:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    return nvcc if os.path.exists(nvcc) else "nvcc"


def _run(cmd):
    print("[build]", " ".join(cmd), flush=True)
    subprocess.run(cmd, cwd=HERE, check=True)


def build(force: bool = False, verbose: bool = False, out: str = OUT) -> str:
    if not force and out == OUT and not stale():
        return OUT
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    extra = ["-Xptxas", "-v"] if verbose else []
    n = num_plans()
    objs = [os.path.join(OBJ, "sfe_lib.o")] + [os.path.join(OBJ, f"sfe_plan_{i}.o") for i in range(n)]
    jobs = [[nvcc] + NVCC_FLAGS + extra + ["-c", "sfe_lib.cu", "-o", objs[0]]]
    for i in range(n):
        jobs.append([nvcc] + NVCC_FLAGS + extra + [f"-DSFE_PLAN_TU={i}", "-c", "sfe_plans.cu", "-o", objs[1 + i]])
    dev = os.environ.get("SFE_DEV_PLANS")
    if dev:  # development shortcut: the other objects must already exist
        toks = [x.strip() for x in dev.split(",") if x.strip()]
        keep = {int(x) for x in toks if x.isdigit()}
        jobs = [j for k, j in enumerate(jobs) if (k == 0 and "lib" in toks) or (k - 1) in keep or not os.path.exists(objs[k])]
    workers = int(os.environ.get("SFE_BUILD_JOBS", str(os.cpu_count() or 4)))
    with ThreadPoolExecutor(max_workers=max(1, workers)) as ex:
        list(ex.map(_run, jobs))
    _run([nvcc] + NVCC_FLAGS + ["-shared", "-o", out] + objs)
    return out


DRAW_DIR = os.path.join(HERE, "drawext")
DRAW_NAME = "sfe_draw"


def draw_ext_path() -> str:
    return os.path.join(DRAW_DIR, DRAW_NAME + ".so")


def build_draw_ext(force: bool = False) -> str:
    """The host-side draw helper (sfe_draw.cpp: SpecAugment draws on torch's CPU generator in one call): a torch C++
    extension, built in-tree with torch.utils.cpp_extension.  Optional — without it the module layer makes the same
    draws with tensor ops."""
    out, src = draw_ext_path(), os.path.join(HERE, "sfe_draw.cpp")
    if not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    from torch.utils.cpp_extension import load
    os.makedirs(DRAW_DIR, exist_ok=True)
    load(name=DRAW_NAME, sources=[src], build_directory=DRAW_DIR, extra_cflags=["-O2"], verbose=False)
    return out


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    build_draw_ext(force="--force" in sys.argv)
