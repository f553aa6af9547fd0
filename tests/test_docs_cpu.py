"""CPU: every repository path DESIGN.md / INTEGRATION.md / README.md cite in back-ticks exists, and
every entry point include/coopcap.h declares is named in INTEGRATION.md's tables."""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PREFIXES = "profiles|tools|tests|oracle|include|cooperativeimagecaptioning_b200|csrc"


def _expand(path):
    if "{" not in path:
        return [path]
    pre, rest = path.split("{", 1)
    opts, post = rest.split("}", 1)
    return [pre + o + post for o in opts.split(",")]


def test_cited_paths_exist():
    missing = []
    for doc in ("DESIGN.md", "INTEGRATION.md", "README.md"):
        text = open(os.path.join(ROOT, doc)).read()
        for m in sorted(set(re.findall(r"`((?:%s)/[^`\s]+)`" % PREFIXES, text))):
            path = m.split("::")[0]
            if path.startswith("csrc/"):
                path = "cooperativeimagecaptioning_b200/" + path
            for c in _expand(path):
                full = os.path.join(ROOT, c)
                if not (glob.glob(full) or os.path.exists(full)):
                    missing.append((doc, c))
    assert not missing, missing


def test_every_entry_point_is_documented():
    hdr = open(os.path.join(ROOT, "include", "coopcap.h")).read()
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    names = sorted(set(re.findall(r"\b(coopcap_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 30
    # families written with a shared prefix in the tables (coopcap_prof_enable / ..._kinds / ..._report)
    undocumented = [n for n in names if n not in doc]
    assert not undocumented, undocumented
