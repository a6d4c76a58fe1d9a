"""Recipe that vendors the UNMODIFIED reference model into oracle/_ref/ (git-ignored, travels to the GPU box with the snapshot).

TEST / BENCH INFRASTRUCTURE.  The reference is 201 lines of pure Python (src/model_fibinet.py) plus src/utils.py; it needs no
build.  `/root/reference` exists only in the dev container, so `__graft_entry__.build()` runs this there and bench.py's CPU
legs (`cpu_baseline`, `--impl reference`) then time the reference's own module (kind = "reference") on the GPU box's host
cores.  Nothing is copied into the git history; the product never imports oracle/_ref.

    python oracle/make_ref.py            # copies /root/reference/src/*.py byte for byte
"""
from __future__ import annotations

import hashlib
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/src"
DST = os.path.join(HERE, "_ref")
FILES = ("model_fibinet.py", "utils.py")                       # what the CPU arm of bench.py imports
SCRIPTS = ("train_fibinet.py", "Prediction.py", "dataloader.py")  # the reference's own entry points, driven against the swapped module
                                                                  # by tests/test_gpu_entrypoints.py ("Entry-point" row)


def make_ref() -> str | None:
    """Returns the directory holding the vendored files, or None when neither the reference nor an earlier copy exists."""
    if os.path.isdir(SRC):
        os.makedirs(DST, exist_ok=True)
        lines = []
        for f in FILES + SCRIPTS:
            shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
            with open(os.path.join(DST, f), "rb") as fh:
                lines.append(f"{hashlib.sha256(fh.read()).hexdigest()}  {f}")
        with open(os.path.join(DST, "SHA256SUMS"), "w") as fh:
            fh.write("\n".join(lines) + "\n")
    return DST if all(os.path.exists(os.path.join(DST, f)) for f in FILES) else None


def ref_dir() -> str | None:
    return DST if all(os.path.exists(os.path.join(DST, f)) for f in FILES) else None


if __name__ == "__main__":
    print(make_ref())
