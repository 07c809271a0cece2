"""TEST / BENCH INFRASTRUCTURE — stages the UNMODIFIED reference under oracle/_ref/ so that it can travel to the GPU box.

    python -m oracle.stage_ref            (also run by __graft_entry__.build() when /root/reference is present)

The reference is pure Python (2.4 k lines, no build step): "building" it is archiving its files byte for byte into
oracle/_ref/reference_src.zip (+ MANIFEST.json with the sha256 of every member).  oracle/_ref/ is git-ignored — the
sources never enter this repository's history or working tree as files — but not gpurun-ignored, so
`bench.py --impl reference` can time the reference's own code on the GPU box's host cores, where /root/reference does
not exist.  oracle.ref_shims unpacks the archive into a temporary directory at run time and verifies the hashes.
"""
from __future__ import annotations

import hashlib
import json
import os
import tempfile
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
ARCHIVE = os.path.join(DST, "reference_src.zip")
MANIFEST = os.path.join(DST, "MANIFEST.json")
SRC = os.environ.get("NEXTOU_REFERENCE_SRC", "/root/reference")
DIRS = ("network_architecture", "loss", "nnUNetTrainer")


def stage() -> str | None:
    """Archive the reference's python files (no edits).  Returns the archive path, or None when neither the reference
    checkout nor a previously staged archive exists (GPU box: the archive that travelled with the snapshot is used)."""
    if not os.path.isfile(os.path.join(SRC, "network_architecture", "NexToU.py")):
        return ARCHIVE if os.path.isfile(ARCHIVE) and os.path.isfile(MANIFEST) else None
    os.makedirs(DST, exist_ok=True)
    manifest = {}
    with zipfile.ZipFile(ARCHIVE, "w", zipfile.ZIP_DEFLATED) as z:
        for d in DIRS:
            for f in sorted(os.listdir(os.path.join(SRC, d))):
                if f.endswith(".py"):
                    data = open(os.path.join(SRC, d, f), "rb").read()
                    manifest[f"{d}/{f}"] = hashlib.sha256(data).hexdigest()
                    z.writestr(zipfile.ZipInfo(f"{d}/{f}", date_time=(2020, 1, 1, 0, 0, 0)), data)
        if os.path.exists(os.path.join(SRC, "LICENSE")):
            z.writestr(zipfile.ZipInfo("LICENSE", date_time=(2020, 1, 1, 0, 0, 0)), open(os.path.join(SRC, "LICENSE"), "rb").read())
    with open(MANIFEST, "w") as fh:
        json.dump({"source": SRC, "files": manifest}, fh, indent=1, sort_keys=True)
    return ARCHIVE


_UNPACKED = None


def unpack() -> str | None:
    """Extract the staged archive into a fresh temporary directory (once per process), check every member against the
    manifest, and return that directory — or None if nothing is staged / a hash does not match."""
    global _UNPACKED
    if _UNPACKED is not None:
        return _UNPACKED
    if not (os.path.isfile(ARCHIVE) and os.path.isfile(MANIFEST)):
        return None
    manifest = json.load(open(MANIFEST))["files"]
    root = tempfile.mkdtemp(prefix="nextou_ref_")
    with zipfile.ZipFile(ARCHIVE) as z:
        for rel, sha in manifest.items():
            data = z.read(rel)
            if hashlib.sha256(data).hexdigest() != sha:
                return None
            os.makedirs(os.path.dirname(os.path.join(root, rel)), exist_ok=True)
            with open(os.path.join(root, rel), "wb") as fh:
                fh.write(data)
    _UNPACKED = root
    return root


if __name__ == "__main__":
    print(stage())
