"""Recipe that stages the UNMODIFIED reference under the git-ignored ``baseline/_ref/`` (test + benchmark infrastructure).

The reference (Rongjiehuang/FastDiff) is pure Python with no ``setup.py`` -- nothing to ``pip install`` and nothing to compile -- and
``/root/reference`` does not exist on the GPU box.  ``baseline/_ref/`` is listed in ``.gitignore`` (never enters history, never
counted as product source) but NOT in ``.gpurunignore``, so the staged tree travels to the B200 box with the snapshot, where

  * ``bench.py --impl reference`` and the ``cpu_baseline`` leg time the reference's OWN ``sampling_given_noise_schedule`` /
    ``FastDiff.forward`` (``cpu_baseline.kind = "reference"``) instead of the oracle port, and
  * the ``-m gpu`` drop-in tests run the reference's own task / vocoder-registry code against the CUDA path.

What is staged: the Python packages the sampling path and its callers import (``modules/ tasks/ utils/ vocoders/ data_gen/``),
their YAML configs, and the three LJSpeech wavs under ``egs/audios`` (fixtures).  Nothing is edited.  Run by
``__graft_entry__.build()`` whenever ``/root/reference`` is present; a no-op elsewhere.
"""
from __future__ import annotations

import os
import shutil
import sys

SRC = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "baseline", "_ref")
TREES = ("modules", "tasks", "utils", "vocoders", "data_gen", os.path.join("egs", "audios"))
KEEP = (".py", ".yaml", ".yml", ".wav")


def stage(force: bool = False) -> str | None:
    """Copy the trees (by extension) if the source is present; returns the staged root or None."""
    if not os.path.isdir(os.path.join(SRC, "modules", "FastDiff")):
        return DST if os.path.isdir(os.path.join(DST, "modules", "FastDiff")) else None
    n = 0
    for tree in TREES:
        for dirpath, _dirs, files in os.walk(os.path.join(SRC, tree)):
            rel = os.path.relpath(dirpath, SRC)
            for f in files:
                if not f.endswith(KEEP):
                    continue
                s, d = os.path.join(dirpath, f), os.path.join(DST, rel, f)
                if not force and os.path.exists(d) and os.path.getmtime(d) >= os.path.getmtime(s) and os.path.getsize(d) == os.path.getsize(s):
                    continue
                os.makedirs(os.path.dirname(d), exist_ok=True)
                shutil.copy2(s, d)
                n += 1
    with open(os.path.join(DST, "STAGED_FROM"), "w") as fh:
        fh.write("unmodified copy of /root/reference (python, yaml, egs/audios wavs) staged by oracle/stage_reference.py\n")
    return DST


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv))
