"""Recipe for `oracle/_ref/`: the reference's OWN code of the hot path, made importable where /root/reference does not exist (the GPU box).

The reference is pure Python: "building" it means laying its package trees out under oracle/_ref/ next to the import-time
stubs of the three packages it imports but never executes on this path (accelerate, peft, imageio - tests/golden/ref_stubs).  Nothing is
edited.  oracle/_ref/ is git-ignored (reference sources never enter the repository's history) but travels with `gpurun`, like the built .so.

    python tools/make_oracle_ref.py            # in the build container (needs /root/reference); __graft_entry__.build() runs it when possible

Consumers: `bench.py --impl reference` and the `cpu_baseline` leg (kind "reference": SD3Transformer2DModel.forward +
FlowMatchEulerDiscreteSDEScheduler.step in the sd3_5.py:273-304 loop, CPU bf16 autocast) and tests/ - never the product.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
DST = os.path.join(ROOT, "oracle", "_ref")
TREES = (("diffusers/src/diffusers", "diffusers"), ("src/flow_factory", "flow_factory"))
STUBS = os.path.join(ROOT, "tests", "golden", "ref_stubs")


def _git_rev(path: str) -> str:
    try:
        return subprocess.run(["git", "-C", path, "rev-parse", "--short", "HEAD"], capture_output=True, text=True, timeout=20).stdout.strip()
    except Exception:
        return ""


def available() -> bool:
    return os.path.isdir(os.path.join(DST, "diffusers")) and os.path.isdir(os.path.join(DST, "flow_factory"))


def make(force: bool = False) -> bool:
    """Returns True when oracle/_ref is usable afterwards."""
    if not os.path.isdir(os.path.join(REF, "src", "flow_factory")):
        return available()                      # GPU box: use what travelled
    if available() and not force:
        return True
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    keep = lambda d, names: [n for n in names if n == "__pycache__" or n.endswith((".pyc", ".so", ".png", ".jpg", ".gif", ".mp4"))]
    for src, name in TREES:
        shutil.copytree(os.path.join(REF, src), os.path.join(DST, name), ignore=keep)
    for name in os.listdir(STUBS):
        p = os.path.join(STUBS, name)
        if os.path.isdir(p):
            shutil.copytree(p, os.path.join(DST, name), ignore=keep)
    with open(os.path.join(DST, "PROVENANCE.txt"), "w") as f:
        f.write("Unmodified copies made by tools/make_oracle_ref.py (not tracked by git):\n"
                f"  flow_factory  <- {REF}/src/flow_factory        (commit {_git_rev(REF) or "a0b2bc5, as surveyed"})\n"
                f"  diffusers     <- {REF}/diffusers/src/diffusers (commit {_git_rev(os.path.join(REF, 'diffusers')) or "f7fd76a, as surveyed"})\n"
                "  accelerate / peft / imageio (+ dist-info) <- tests/golden/ref_stubs (import-time stubs, this repo)\n")
    return True


def import_path() -> str:
    return DST


if __name__ == "__main__":
    ok = make(force="--force" in sys.argv)
    n = sum(len(fs) for _, _, fs in os.walk(DST)) if ok else 0
    print(f"oracle/_ref: {'ok' if ok else 'unavailable'} ({n} files)")
