"""TEST INFRASTRUCTURE / CPU BASELINE — stage the reference's hot-path model files for the GPU box.

``/root/reference`` exists only in the build container.  This recipe copies, UNMODIFIED and from where they lie, the Python
files the five reference adapter modules need at import time (the model packages that define ``ExpertAdapter`` /
``MoEAdapter`` and their local imports) into ``oracle/_ref/`` — git-ignored, so no reference source ever enters the history,
but not gpurun-ignored, so the copy travels to the GPU box like a built ``.so``.  There ``oracle/ref_loader.py`` imports the
classes from the staged tree, and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` leg times the reference's own classes
(``kind: "reference"``) instead of the oracle port.  ``__graft_entry__.build()`` runs it when ``/root/reference`` is present.

    python -m oracle.stage_reference          # -> oracle/_ref/AVMOE/...
"""
from __future__ import annotations

import glob
import os
import shutil

SRC_ROOT = "/root/reference"
DST_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")

# model packages that hold the hot-path classes: every top-level *.py of each directory
PACKAGE_DIRS = [
    "AVMOE/AVE/nets",                         # net_trans_v3.py:296-487
    "AVMOE/AVQA/net_grd_avst",                # net_avst_v2.py:215-399
    "AVMOE/AVVP/nets",                        # mgn.py:39-224
    "AVMOE/AVS/avs_scripts/avs_s4/model",     # PVT_AVSModel_v2.py:90-318
    "AVMOE/AVS/avs_scripts/avs_ms3/model",    # PVT_AVSModel_v2.py:282-501
]


def stage(src_root: str = SRC_ROOT, dst_root: str = DST_ROOT) -> int:
    """Copy the files; returns how many were staged (0 when the reference tree is absent)."""
    if not os.path.isdir(os.path.join(src_root, "AVMOE")):
        return 0
    n = 0
    for rel in PACKAGE_DIRS:
        dst_dir = os.path.join(dst_root, rel)
        os.makedirs(dst_dir, exist_ok=True)
        for f in sorted(glob.glob(os.path.join(src_root, rel, "*.py"))):
            shutil.copyfile(f, os.path.join(dst_dir, os.path.basename(f)))
            n += 1
    return n


if __name__ == "__main__":
    print(f"staged {stage()} files under {DST_ROOT}")
