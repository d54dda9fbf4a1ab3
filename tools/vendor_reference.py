#!/usr/bin/env python
"""Copy the UNMODIFIED reference modules into baseline/_ref/ (git-ignored, NOT gpurun-ignored) so that the reference
arms of bench.py and the drop-in test can run them on the GPU box, where /root/reference does not exist.

The reference (KangJialiang/ViSpec d0240c87) is 100 % Python without setup.py / pyproject, so there is nothing
`pip install --target baseline/_ref /root/reference` could build: the "install" is this copy of `vispec/` (model package
+ the evaluation script the drop-in test drives).  Nothing is edited; baseline/_ref/VENDORED.json records the sha256 of
every file so a later check can prove it.  Called by __graft_entry__.build() whenever /root/reference is present.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = ["vispec/__init__.py", "vispec/evaluation/__init__.py", "vispec/evaluation/gen_spec_answer_coco_caption.py",
         "vispec/evaluation/gen_baseline_answer_coco_caption.py", "vispec/evaluation/coco_caption_prompt.py",
         "vispec/evaluation/speed.py"]


def _sha(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def vendor(src: str = SRC, dst: str = DST) -> dict:
    """Returns the manifest {relative path: sha256}.  No-op (returns the existing manifest) when `src` is absent."""
    man_path = os.path.join(dst, "VENDORED.json")
    if not os.path.isdir(os.path.join(src, "vispec", "model")):
        if os.path.exists(man_path):
            with open(man_path) as f:
                return json.load(f)
        return {}
    files = list(FILES)
    for name in sorted(os.listdir(os.path.join(src, "vispec", "model"))):
        if name.endswith(".py"):
            files.append("vispec/model/" + name)
    for name in sorted(os.listdir(os.path.join(src, "vispec", "train"))):
        if name.endswith(".json"):                       # draft-head configs (llava_1.6_7B_config.json, ...)
            files.append("vispec/train/" + name)
    manifest = {}
    for rel in files:
        s = os.path.join(src, rel)
        if not os.path.exists(s):
            continue
        d = os.path.join(dst, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        manifest[rel] = _sha(d)
    with open(man_path, "w") as f:
        json.dump({"source": src, "files": manifest}, f, indent=1, sort_keys=True)
    return {"source": src, "files": manifest}


def verify(dst: str = DST) -> bool:
    """True when every vendored file still has the recorded hash (i.e. nobody edited the copy)."""
    man_path = os.path.join(dst, "VENDORED.json")
    if not os.path.exists(man_path):
        return False
    with open(man_path) as f:
        man = json.load(f)
    return all(os.path.exists(os.path.join(dst, rel)) and _sha(os.path.join(dst, rel)) == h
               for rel, h in man.get("files", {}).items())


if __name__ == "__main__":
    m = vendor()
    print(f"vendored {len(m.get('files', {}))} files into {DST}; verify: {verify()}", file=sys.stderr)
