"""Recipe for `oracle/_ref/`: an UNMODIFIED copy of the reference's Python sources for the hot path.

TEST / MEASUREMENT INFRASTRUCTURE (see oracle/__init__.py).

The reference is pure Python, so "building" it means making it importable where `/root/reference` does not exist: the
GPU box.  `python -m oracle.make_ref` (also called by `__graft_entry__.build()` whenever `/root/reference` is present)
copies, byte for byte,

    /root/reference/remfx/*.py            -> oracle/_ref/remfx/
    /root/reference/umx/openunmix/*.py    -> oracle/_ref/umx/openunmix/
    /root/reference/cfg/**                -> oracle/_ref/cfg/
    /root/reference/example.wav           -> oracle/_ref/example.wav

`oracle/_ref/` is listed in `.gitignore` (never committed: no reference source enters the history) but NOT in
`.gpurunignore`, so -- like the built `.so` -- it travels with the snapshot.  `oracle/refshim.py` imports the modules
from `/root/reference` when that exists and from `oracle/_ref/` otherwise; `bench.py --impl reference` and the
GPU-eager bar time exactly these unchanged modules.  A manifest with the SHA-256 of every copied file is written next
to them so a reader can check that nothing was edited.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC = os.environ.get("REMFX_REFERENCE", "/root/reference")

TREES = (("remfx", (".py",)), (os.path.join("umx", "openunmix"), (".py",)), ("cfg", (".yaml", ".yml")))
FILES = ("example.wav", "LICENSE", os.path.join("umx", "LICENSE"))


def make(verbose: bool = False) -> str | None:
    if not os.path.isdir(os.path.join(SRC, "remfx")):
        return None  # not in the build container: keep whatever copy travelled with the snapshot
    manifest = {}

    def put(rel: str) -> None:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        with open(d, "rb") as fh:
            manifest[rel] = hashlib.sha256(fh.read()).hexdigest()

    if os.path.isdir(DST):
        shutil.rmtree(DST)
    for tree, exts in TREES:
        for root, _, names in os.walk(os.path.join(SRC, tree)):
            for n in sorted(names):
                if n.endswith(exts):
                    put(os.path.relpath(os.path.join(root, n), SRC))
    for f in FILES:
        if os.path.exists(os.path.join(SRC, f)):
            put(f)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC, "files": manifest}, fh, indent=1, sort_keys=True)
    if verbose:
        print(f"oracle/_ref: {len(manifest)} files copied unmodified from {SRC}")
    return DST


if __name__ == "__main__":
    out = make(verbose=True)
    if out is None:
        print(f"{SRC} not present: nothing copied", file=sys.stderr)
