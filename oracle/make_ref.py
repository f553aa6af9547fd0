"""TEST / BENCH INFRASTRUCTURE ONLY -- recipe for `baseline/_ref/`: the REAL reference's hot-path
modules, made importable on torch 2.x by the mechanical patch of SURVEY.md §8(c).

    python -m oracle.make_ref            # /root/reference  ->  <repo>/baseline/_ref/

`baseline/_ref/` is git-ignored (reference sources never enter this repository's history) but
travels to the GPU box with the snapshot, where `bench.py --impl reference` times it on the host
cores.  The reference tree is read, never written.  What is taken: `models/*.py`, `misc/utils.py`,
`misc/__init__.py`, `optimizer.py`, `misc/rewards.py` and the pure-Python CIDEr-D scorer it
imports (`cider/pyciderevalcap/ciderD/*.py`).  The patch (regular expressions on the text, file by
file; nothing else is edited):

  R1  `.data[0]` -> `.item()`                      0-dim indexing was removed after torch 0.4
  R2  drop `import skimage`, `skimage.io`, `skimage.transform`, `scipy.misc.imresize`
                                                   (misc/utils.py:8-11; image loading, unused here)
  R3  `from cider...cider_diff.cider import Cider` -> `Cider = None`
                                                   (AlternatingJointModel.py:53; its class body
                                                   loads a blob that is not in the tree, the name
                                                   is never used)
  R4  `list(sorted_lens.data)` -> `sorted_lens.cpu().tolist()`   (VSEFCModel.py:108; CUDA runs only)
  R5  Python-2 idioms of the CIDEr-D scorer: `xrange`, `dict.iteritems()`, `cPickle`
                                                   (ciderD_scorer.py; only reached with
                                                   cider_optimization = 1)

`build()` in __graft_entry__.py calls `make()` when the reference tree is present.
"""
from __future__ import annotations

import hashlib
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE_ROOT = os.environ.get("COOPCAP_REFERENCE_ROOT", "/root/reference")
DEST = os.path.join(ROOT, "baseline", "_ref")

FILES = [
    "models/__init__.py", "models/AttModel.py", "models/FCModel.py", "models/VSEFCModel.py",
    "models/AlternatingJointModel.py", "models/gumbel.py", "models/gumbel_softmax.py",
    "models/multinomial.py", "models/multinomial_soft.py",
    "misc/__init__.py", "misc/utils.py", "misc/rewards.py", "optimizer.py",
    "cider/pyciderevalcap/__init__.py", "cider/pyciderevalcap/ciderD/__init__.py",
    "cider/pyciderevalcap/ciderD/ciderD.py", "cider/pyciderevalcap/ciderD/ciderD_scorer.py",
]


def patch(rel: str, src: str) -> str:
    src = re.sub(r"\.data\[\s*0\s*\]", ".item()", src)                                   # R1
    if rel == "misc/utils.py":                                                           # R2
        src = re.sub(r"^import skimage.*$", "", src, flags=re.M)
        src = re.sub(r"^from scipy\.misc import imresize.*$", "", src, flags=re.M)
    if rel == "models/AlternatingJointModel.py":                                         # R3
        src = re.sub(r"^from cider\.pyciderevalcap\.cider_diff\.cider import Cider.*$",
                     "Cider = None", src, flags=re.M)
    if rel == "models/VSEFCModel.py":                                                    # R4
        src = src.replace("list(sorted_lens.data)", "sorted_lens.cpu().tolist()")
    if rel.startswith("cider/"):                                                         # R5
        src = re.sub(r"\bxrange\(", "range(", src)
        src = re.sub(r"\.iteritems\(\)", ".items()", src)
        src = re.sub(r"^import cPickle$", "import pickle as cPickle", src, flags=re.M)
        src = re.sub(r"^from six\.moves import cPickle$", "import pickle as cPickle", src, flags=re.M)
    return src


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "AttModel.py"))


def make(dest: str = DEST) -> str:
    if not available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    manifest = {}
    for rel in FILES:
        with open(os.path.join(REFERENCE_ROOT, rel), "r") as f:
            src = f.read()
        out = patch(rel, src)
        path = os.path.join(dest, rel)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as f:
            f.write(out)
        manifest[rel] = dict(source_sha1=hashlib.sha1(src.encode()).hexdigest(),
                             patched=(out != src))
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump(dict(reference_root=REFERENCE_ROOT, rules="SURVEY.md §8(c) R1-R5 (oracle/make_ref.py)",
                       files=manifest), f, indent=1, sort_keys=True)
    return dest


if __name__ == "__main__":
    print(make(sys.argv[1] if len(sys.argv) > 1 else DEST))
