"""TEST INFRASTRUCTURE ONLY -- import the *real* reference from /root/reference, in memory.

The reference (torch 0.4.1-era Python) does not import on torch 2.x as it lies.  This loader
serves `models.*` and `misc.*` straight from the read-only reference tree through an import hook
and applies the mechanical patch of SURVEY.md §8(c) to the *source text in memory* -- nothing is
copied into this repository and nothing is written to /root/reference:

  R1  `.data[0]`  ->  `.item()`                       (0-dim indexing removed after torch 0.4)
  R2  drop the unused skimage / scipy.misc imports    (misc/utils.py:8-11)
  R3  drop the cider_diff import                      (models/AlternatingJointModel.py:53; the
      class body needs a large blob that is not in the tree; the name is never used).
      `misc.rewards` and the pure-Python CIDEr-D scorer it imports load unmodified
      (`<reference>/cider` is put on sys.path, as misc/rewards.py:13-16 does itself)

It exists so that (a) the oracle restatement in oracle/*.py can be validated against the code it
restates and (b) tests/golden/make_golden.py can produce golden vectors.  The reference tree is
only present in the development container; nothing that runs on the GPU box imports this module.
"""
from __future__ import annotations

import importlib.abc
import importlib.util
import os
import re
import sys
import types

REFERENCE_ROOT = os.environ.get("COOPCAP_REFERENCE_ROOT", "/root/reference")

_PKGS = ("models", "misc")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "AttModel.py"))


def _patch(modname: str, src: str) -> str:
    src = re.sub(r"\.data\[\s*0\s*\]", ".item()", src)                       # R1
    if modname == "misc.utils":                                              # R2
        src = re.sub(r"^import skimage.*$", "", src, flags=re.M)
        src = re.sub(r"^from scipy\.misc import imresize.*$", "", src, flags=re.M)
    if modname == "models.AlternatingJointModel":                            # R3
        src = re.sub(r"^from cider\.pyciderevalcap\.cider_diff\.cider import Cider.*$",
                     "Cider = None", src, flags=re.M)
    return src


class _RefLoader(importlib.abc.Loader):
    def __init__(self, fullname, path, is_pkg):
        self.fullname, self.path, self.is_pkg = fullname, path, is_pkg

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        with open(self.path, "r") as f:
            src = _patch(self.fullname, f.read())
        module.__dict__.setdefault("__file__", self.path)    # misc/rewards.py:15-16 reads it
        exec(compile(src, self.path, "exec"), module.__dict__)


class _RefFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        parts = fullname.split(".")
        if parts[0] not in _PKGS or len(parts) > 2:
            return None
        base = os.path.join(REFERENCE_ROOT, *parts)
        if os.path.isdir(base):
            init = os.path.join(base, "__init__.py")
            spec = importlib.util.spec_from_loader(fullname, _RefLoader(fullname, init, True),
                                                   origin=init, is_package=True)
            spec.submodule_search_locations = [base]
            return spec
        if os.path.isfile(base + ".py"):
            return importlib.util.spec_from_loader(
                fullname, _RefLoader(fullname, base + ".py", False), origin=base + ".py")
        return None


_installed = False


def load_reference() -> types.ModuleType:
    """Return the reference's `models` package (with `models.AlternatingJointModel`, `setup`, ...)."""
    global _installed
    if not available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    if not _installed:
        for k in list(sys.modules):
            if k.split(".")[0] in _PKGS:
                raise RuntimeError(f"module {k!r} already imported; cannot hook the reference")
        sys.meta_path.insert(0, _RefFinder())
        sys.path.append(os.path.join(REFERENCE_ROOT, "cider"))
        _installed = True
    import models  # noqa: F401  (served by _RefFinder)
    return sys.modules["models"]


def reference_opt(**overrides):
    """An argparse-like Namespace with the reference defaults that the hot path reads
    (opts.py:36-68,94-103,192-234 and the loader-provided vocab_size / seq_length)."""
    import argparse
    d = dict(
        vocab_size=9487, seq_length=16, input_encoding_size=512, rnn_size=512, num_layers=1,
        drop_prob_lm=0.5, fc_feat_size=2048, att_feat_size=2048, att_hid_size=512, use_bn=0,
        decoding_constraint=0, retrieval_reward="gumbel", gumbel_temp=1.0, multinomial_temp=1.0,
        prob_gumbel_softmax=0.25, prob_multinomial_soft=0.25,
        caption_model="att2in2", vse_model="fc", vse_embed_size=1024, vse_no_imgnorm=0,
        vse_use_abs=0, vse_num_layers=1, vse_rnn_type="gru", vse_pool_type="last",
        vse_margin=0.2, vse_measure="cosine", vse_max_violation=1, vse_loss_type="contrastive",
        share_embed=0, phase=None, batch_size=10, vse_loss_weight=0.0, caption_loss_weight=0.0,
        retrieval_reward_weight=0.01, reinforce_baseline_type="gt", only_one_retrieval="off",
        cider_optimization=0, use_gen_cider_scores=0, is_alternating=0, alternating_turn=None,
        continue_from_existing_models=True, start_from=None, initialize_retrieval=None,
        id="", grad_clip=0.1, learning_rate=5e-4, weight_decay=0.0,
    )
    d.update(overrides)
    return argparse.Namespace(**d)
