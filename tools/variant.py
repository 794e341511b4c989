"""Development helper: build tools/ab/libsfe_<name>.so with extra -D flags applied to chosen plan objects only.

    python tools/variant.py name "-DSFE_ABLATE=1" 12[,13]

The other objects are reused from speechain_b200/csrc/build (run the normal build first)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from speechain_b200.csrc import build as B  # noqa: E402

name, flags, plans = sys.argv[1], sys.argv[2].split(), [int(x) for x in sys.argv[3].split(",")]
out_dir = os.path.join(ROOT, "tools", "ab")
os.makedirs(out_dir, exist_ok=True)
n = B.num_plans()
objs = [os.path.join(B.OBJ, "sfe_lib.o")] + [os.path.join(B.OBJ, f"sfe_plan_{i}.o") for i in range(n)]
procs = []
for i in plans:
    o = os.path.join(out_dir, f"{name}_plan_{i}.o")
    objs[1 + i] = o
    cmd = [B._nvcc()] + B.NVCC_FLAGS + flags + [f"-DSFE_PLAN_TU={i}", "-c", "sfe_plans.cu", "-o", o]
    procs.append(subprocess.Popen(cmd, cwd=B.HERE))
for p in procs:
    assert p.wait() == 0
out = os.path.join(out_dir, f"libsfe_{name}.so")
subprocess.run([B._nvcc()] + B.NVCC_FLAGS + ["-shared", "-o", out] + objs, cwd=B.HERE, check=True)
print(out)
