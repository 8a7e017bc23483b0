"""Stages the UNMODIFIED reference's Python packages (asr/, utils/, lm/) under baseline/_ref/emoASR so that they travel
to the GPU box with the gpurun snapshot (baseline/_ref is git-ignored: nothing of the reference enters the
history).  tests/test_gpu_reference_train.py runs the reference's own asr/train_asr.py from there through the
drop-in launcher.  Called by __graft_entry__.build() when /root/reference is present."""
import os
import shutil
import sys

SRC = os.environ.get("EMOASR_REFERENCE", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "emoASR")


def stage(src=SRC, dst=DST):
    if not os.path.isdir(os.path.join(src, "asr")):
        return None
    for pkg in ("asr", "utils", "lm"):
        shutil.copytree(os.path.join(src, pkg), os.path.join(dst, pkg), dirs_exist_ok=True,
                        ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    return dst


if __name__ == "__main__":
    print(stage() or f"no reference at {SRC}", file=sys.stderr)
