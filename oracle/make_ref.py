"""TEST INFRASTRUCTURE ONLY — recipe that makes the *unmodified* reference importable on the GPU box.

    python -m oracle.make_ref

The reference (whisperX 3.1.1) is pure Python: there is nothing to compile.  What bench.py's
`--impl reference` arm and the `cpu_baseline` leg need is the reference's own `whisperx/alignment.py`
(and the three modules it imports relatively: audio.py, utils.py, types.py).  /root/reference does not
exist on the GPU box, so at build() time — in the build container, where it does exist — those four
files are copied byte for byte into oracle/_ref/whisperx/.  oracle/_ref/ is git-ignored (no reference
source enters the history) but not gpurun-ignored, so the copy travels with the snapshot like a built .so.
oracle/ref_import.py imports from /root/reference when present, else from this copy.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("WXB_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ("alignment.py", "audio.py", "utils.py", "types.py", "diarize.py")


def make(verbose: bool = False) -> bool:
    """Copy the files; returns False (and leaves any earlier copy alone) when the reference is absent."""
    src_pkg = os.path.join(SRC, "whisperx")
    if not os.path.isfile(os.path.join(src_pkg, "alignment.py")):
        return False
    dst_pkg = os.path.join(DST, "whisperx")
    os.makedirs(dst_pkg, exist_ok=True)
    manifest = {}
    for name in FILES:
        a, b = os.path.join(src_pkg, name), os.path.join(dst_pkg, name)
        if not os.path.isfile(a):
            continue
        shutil.copyfile(a, b)
        manifest[name] = hashlib.sha256(open(b, "rb").read()).hexdigest()
    for name in ("LICENSE",):
        if os.path.isfile(os.path.join(SRC, name)):
            shutil.copyfile(os.path.join(SRC, name), os.path.join(DST, name))
    json.dump({"source": SRC, "sha256": manifest}, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    if verbose:
        print(f"oracle/_ref: {len(manifest)} reference files copied from {src_pkg}")
    return True


if __name__ == "__main__":
    print("copied" if make(verbose=True) else f"{SRC} not present; nothing copied")
