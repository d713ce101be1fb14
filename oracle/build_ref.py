"""Recipe for ``oracle/_ref/``: the UNMODIFIED reference hot path, taken from where it lies.

TEST / BENCH INFRASTRUCTURE, NOT PRODUCT CODE.  The reference (hugofloresgarcia/ddsp_pytorch) is pure Python, so
"building" it means placing its own files where the GPU box can import them: ``/root/reference`` does not exist
there, ``oracle/_ref/`` travels with the repo snapshot (git-ignored, NOT gpurun-ignored -- like the built ``.so``
files).  Nothing is copied into the git history; this script is the only committed part.

    python oracle/build_ref.py          # in the build container (needs /root/reference)

Files taken, byte for byte (sha256 recorded in ``oracle/_ref/MANIFEST.json``):
    ddsp/__init__.py  ddsp/core.py  ddsp/utils.py  ddsp/models/__init__.py  ddsp/models/modules.py
    ddsp/models/decoder.py  ddsp/models/encoder.py
``ddsp/data.py`` and ``ddsp/preprocess.py`` are not on the path (out of scope) and are not taken; the
reference's ``ddsp/__init__.py`` imports them, so ``oracle/ref_step.import_reference`` provides empty stubs for
them next to the stubs for librosa / crepe / matplotlib / pytorch_lightning (none is touched by the hot path,
SURVEY 8c).
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("DDSP_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
FILES = ["ddsp/__init__.py", "ddsp/core.py", "ddsp/utils.py", "ddsp/models/__init__.py", "ddsp/models/modules.py",
         "ddsp/models/decoder.py", "ddsp/models/encoder.py"]


def build(force: bool = False) -> bool:
    """Returns True when oracle/_ref is in place (freshly made or already there)."""
    manifest = os.path.join(OUT, "MANIFEST.json")
    if not os.path.isdir(REF):
        return os.path.exists(manifest)            # GPU box: use what travelled
    if os.path.exists(manifest) and not force:
        with open(manifest) as f:
            have = json.load(f)["sha256"]
        if all(os.path.exists(os.path.join(OUT, p)) and
               hashlib.sha256(open(os.path.join(REF, p), "rb").read()).hexdigest() == have.get(p) for p in FILES):
            return True
    sums = {}
    for p in FILES:
        dst = os.path.join(OUT, p)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF, p), dst)
        sums[p] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    with open(manifest, "w") as f:
        json.dump({"source": "hugofloresgarcia/ddsp_pytorch at " + REF, "unmodified": True, "sha256": sums}, f, indent=1)
    print(f"[oracle/build_ref] {len(FILES)} reference files -> {OUT}")
    return True


if __name__ == "__main__":
    ok = build("--force" in sys.argv)
    sys.exit(0 if ok else 1)
