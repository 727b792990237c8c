#!/usr/bin/env python
"""Install the UNMODIFIED reference files of the hot path into git-ignored `baseline/_ref/`.

The reference is not a package (no setup.py / pyproject.toml), so there is nothing for pip to build; the two
files of the path — `src/utils/loss.py` (batch_NN_loss, :40-76) and `src/networks/PointNetCls.py` (the point-cloud
discriminator, :11-214) — are copied byte for byte, together with the licence, so that `bench.py --impl reference`
and its `gpu_eager_baseline` leg can run the reference's OWN code on the GPU box, where /root/reference does not
exist.  `baseline/_ref/` is listed in .gitignore (reference sources never enter this repository's history) and
NOT in .gpurunignore (it travels with the snapshot like the built .so files).  Run by __graft_entry__.build()
whenever /root/reference is present.
"""
from __future__ import annotations

import hashlib
import json
import shutil
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
DEST = HERE / "_ref"
FILES = ("src/utils/loss.py", "src/networks/PointNetCls.py", "LICENSE")


def install(reference_root: str = "/root/reference") -> bool:
    root = Path(reference_root)
    if not root.is_dir():
        return False
    manifest = {}
    for rel in FILES:
        src = root / rel
        dst = DEST / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(dst.read_bytes()).hexdigest()
    (DEST / "MANIFEST.json").write_text(json.dumps({"source": str(root), "sha256": manifest}, indent=1))
    return True


if __name__ == "__main__":
    ok = install(*sys.argv[1:2])
    print("installed baseline/_ref" if ok else "reference tree not found; baseline/_ref left as is")
