"""Same surface as the reference's `models` package (models/__init__.py:11-50):
`setup(opt, model_name, model_type)`, `load(model, opt, iteration)`, `AlternatingJointModel`."""
from __future__ import annotations

import os

import torch

from .AttModel import Att2in2Model, AttModel  # noqa: F401
from .VSEFCModel import VSEFCModel  # noqa: F401

__all__ = ["setup", "load", "AlternatingJointModel"]


def setup(opt, model_name, model_type="caption_model"):
    """models/__init__.py:14-33."""
    if model_type == "caption_model":
        if model_name == "att2in2":
            return Att2in2Model(opt)
        if model_name == "fc":
            raise Exception("Caption model 'fc' (FCModel placeholder speaker) is outside the "
                            "B200 hot path; use 'att2in2'")
        raise Exception("Caption model not supported: {}".format(model_name))
    if model_type == "vse_model":
        if model_name == "fc":
            return VSEFCModel(opt)
        raise Exception("VSE model not supported: {}".format(model_name))
    raise Exception("model_type not supported: {}".format(model_type))


def load_state_dict(model, state_dict):
    """Tolerant loader of misc/utils.py:89-107: copy matching names; shape mismatches are
    flattened and copied up to the common length."""
    own = model.state_dict()
    for k, v in state_dict.items():
        if k not in own:
            continue
        dst = own[k]
        if dst.shape == v.shape:
            dst.copy_(v)
        else:
            n = min(dst.numel(), v.numel())
            dst.view(-1)[:n].copy_(v.reshape(-1)[:n])


def load(model, opt, iteration=None):
    """models/__init__.py:35-50."""
    start = vars(opt).get("start_from", None)
    if start is not None:
        assert os.path.isdir(start), " %s must be a a path" % start
        assert os.path.isfile(os.path.join(start, "infos_" + opt.id + ".pkl")), \
            "infos.pkl file does not exist in path %s" % start
        name = "model-" + iteration + ".pth" if iteration else "model.pth"
        load_state_dict(model, torch.load(os.path.join(start, name), map_location="cpu"))


from .AlternatingJointModel import AlternatingJointModel  # noqa: E402,F401
