#!/usr/bin/env python
"""Stages the reference's OWN hot-path modules under oracle/_ref so that they travel to the GPU box.

TEST / BASELINE INFRASTRUCTURE ONLY.  Nothing in ``speechain_b200/`` may import this or what it stages.

    python oracle/build_ref.py            # in the build container, where /root/reference exists

The acoustic frontend of bagustris/speechain is pure Python on torch + torchaudio; the seven files below import
nothing else of the package, so "building" the reference for this path is staging them — file for
file, unmodified — in a directory that is git-ignored (reference sources never enter this repository's history) but
NOT gpurun-ignored (like libsfe.so it ships with the snapshot).  ``bench.py --impl reference`` and the
``cpu_baseline`` leg then time these modules themselves (``kind: "reference"``, through oracle/ref_modules.py) and
fall back to the bit-identical port (oracle/ref_port_torch.py, ``kind: "port"``) only where the staging is absent.
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SFE_REFERENCE_ROOT", "/root/reference")
FILES = [
    "speechain/module/abs.py",
    "speechain/module/frontend/speech2mel.py",
    "speechain/module/frontend/speech2linear.py",
    "speechain/module/frontend/linear2mel.py",
    "speechain/module/frontend/delta_feat.py",
    "speechain/module/augment/specaug.py",
    "speechain/module/norm/feat_norm.py",
]


def build(quiet: bool = False) -> bool:
    """Returns True when oracle/_ref holds the reference modules afterwards."""
    dst_root = os.path.join(HERE, "_ref")
    if not os.path.isdir(os.path.join(REF, "speechain", "module")):
        return os.path.exists(os.path.join(dst_root, "MANIFEST"))
    lines = []
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(dst_root, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        lines.append(f"{hashlib.sha256(open(src, 'rb').read()).hexdigest()}  {rel}")
    # package markers (empty: the reference's own __init__ files pull in the whole toolkit)
    for d in ["speechain", "speechain/module", "speechain/module/frontend", "speechain/module/augment",
              "speechain/module/norm"]:
        open(os.path.join(dst_root, d, "__init__.py"), "w").close()
    with open(os.path.join(dst_root, "MANIFEST"), "w") as f:
        f.write("\n".join(lines) + "\n")
    if not quiet:
        print(f"staged {len(FILES)} reference modules under {dst_root}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
