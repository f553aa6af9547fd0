"""CPU: the C-ABI library builds, loads, exports every symbol include/coopcap.h declares and agrees
with the header-generated ctypes structs.  No compute calls (there is no GPU here)."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from cooperativeimagecaptioning_b200 import _lib, build
    build.build_library()
    lib = _lib.load()                       # raises if a symbol is missing or a struct size differs
    with open(os.path.join(ROOT, "include", "coopcap.h")) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    declared = set(re.findall(r"\b(coopcap_\w+)\s*\(", src))
    assert declared == set(_lib.FUNCTIONS), declared ^ set(_lib.FUNCTIONS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.coopcap_version() == 200
    for sname, sid in _lib._SIZEOF_IDS.items():
        assert C.sizeof(_lib.STRUCTS[sname]) == lib.coopcap_sizeof(sid)
    assert lib.coopcap_sizeof(99) == -1
    assert lib.coopcap_launch_count() == 0


def test_state_dict_matches_reference_names_and_shapes():
    """checkpoint keys / shapes are part of the drop-in contract."""
    import cooperativeimagecaptioning_b200.models as models
    from oracle.ref_loader import reference_opt
    from oracle import synth
    m = models.AlternatingJointModel(reference_opt())
    sd = m.state_dict()
    want = {"caption_generator." + k: v.shape for k, v in synth.speaker_params(synth.Dims()).items()}
    want.update({"vse." + k: v.shape for k, v in synth.listener_params(synth.Dims()).items()})
    assert {k: v.shape for k, v in sd.items()} == want
    assert sum(p.numel() for p in m.caption_generator.parameters()) == 14452497
    assert sum(p.numel() for p in m.vse.parameters()) == 11681280
    with pytest.raises(Exception):
        models.setup(reference_opt(), "topdown", "caption_model")
    with pytest.raises(Exception):
        models.setup(reference_opt(), "att2in2", "vse_model")


def test_no_cpu_fallback():
    """The product path must fail loudly without CUDA tensors."""
    import cooperativeimagecaptioning_b200.models as models
    from cooperativeimagecaptioning_b200._lib import CoopcapError
    from cooperativeimagecaptioning_b200 import optimizer as OPT
    from oracle.ref_loader import reference_opt
    opt = reference_opt()
    m = models.AlternatingJointModel(opt)
    fc, att = torch.zeros(2, 2048), torch.zeros(2, 4, 2048)
    with pytest.raises(CoopcapError):
        m.sample(fc, att, None, {"sample_max": 1})
    with pytest.raises(CoopcapError):
        m.vse(fc, att, torch.zeros(2, 3, dtype=torch.long), torch.ones(2, 3))
    o = OPT.define_optimizer(m.vse, opt)
    with pytest.raises(CoopcapError):
        o.step()


def test_out_of_scope_options_raise():
    import cooperativeimagecaptioning_b200.models as models
    from oracle.ref_loader import reference_opt
    with pytest.raises(NotImplementedError):
        models.setup(reference_opt(use_bn=1), "att2in2", "caption_model")
    with pytest.raises(NotImplementedError):
        models.setup(reference_opt(vse_loss_type="pair"), "fc", "vse_model")
    with pytest.raises(ValueError):
        models.setup(reference_opt(vse_pool_type="median"), "fc", "vse_model")
    # the listener's pooling / abs / hinge options are on the path
    vse = models.setup(reference_opt(vse_pool_type="mean", vse_use_abs=1, vse_max_violation=0), "fc",
                       "vse_model")
    assert vse._variant() == dict(pool_type="mean", use_abs=True, max_violation=False)
