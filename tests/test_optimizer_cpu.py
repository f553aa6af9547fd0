"""CPU: the optimizer boundary train.py drives (reference optimizer.py:9-47,149-242, train.py:22-23,
395,457): every name exists, the nested dictionaries / file names match, a reference-format
`torch.optim.Adam.state_dict()` round-trips through FlatAdam, and the shared embedding of
--share_embed is aliased (not orphaned) by the second agent's optimizer."""
import argparse
import os

import pytest
import torch

from cooperativeimagecaptioning_b200 import optimizer as OPT
import cooperativeimagecaptioning_b200.models as models


def test_names_train_py_imports_exist():
    # train.py:16,22-23: `import models`; `from optimizer import load_optimizer, save_optimizer,
    # zeroing_optimizer, update_optimizer`; misc.utils.set_lr / clip_gradient act on the result
    for name in ("load_optimizer", "save_optimizer", "zeroing_optimizer", "update_optimizer",
                 "load_optimizer_path", "define_optimizer", "load_state_dict",
                 "load_optimizer_from_checkpoint", "define_speaker_optimizer_joint_training",
                 "define_listener_optimizer_joint_training", "define_pretraining_listener_optimizer",
                 "define_pretraining_speaker_optimizer", "define_only_speaker_optimizer",
                 "set_lr", "clip_gradient"):
        assert callable(getattr(OPT, name)), name
    for name in ("setup", "load", "AlternatingJointModel"):
        assert hasattr(models, name), name


class _Agents(torch.nn.Module):
    """Stand-in with the two attributes load_optimizer reads (model.caption_generator / model.vse)."""

    def __init__(self, share):
        super().__init__()
        torch.manual_seed(0)
        self.vse = torch.nn.ModuleDict(dict(embed=torch.nn.Embedding(11, 4), fc=torch.nn.Linear(4, 3)))
        self.caption_generator = torch.nn.ModuleDict(
            dict(embed=torch.nn.Embedding(11, 4), out=torch.nn.Linear(4, 5)))
        if share:
            self.caption_generator["embed"] = self.vse["embed"]


def _opt(**kw):
    d = dict(is_alternating=1, alternating_turn=["speaker", "listener"], retrieval_reward="gumbel",
             start_from=None, share_embed=0, learning_rate=5e-4, weight_decay=0.0, grad_clip=0.1,
             phase=None, checkpoint_path=None, speaker_stage_2_optimizer_path=None,
             initialize_retrieval=None)
    d.update(kw)
    return argparse.Namespace(**d)


def _fake_moments(optimizer, seed):
    g = torch.Generator().manual_seed(seed)
    optimizer.exp_avg.copy_(torch.randn(optimizer.numel, generator=g))
    optimizer.exp_avg_sq.copy_(torch.rand(optimizer.numel, generator=g))
    optimizer.step_count = 7


def _same_moments(a, b):
    """Per-parameter moment segments are equal (the 16-byte padding lanes are not checkpointed)."""
    for p, oa, ob in zip(a.params, a.offsets, b.offsets):
        n = p.numel()
        if not (torch.equal(a.exp_avg[oa:oa + n], b.exp_avg[ob:ob + n]) and
                torch.equal(a.exp_avg_sq[oa:oa + n], b.exp_avg_sq[ob:ob + n])):
            return False
    return a.step_count == b.step_count


def test_torch_adam_state_dict_round_trips():
    torch.manual_seed(1)
    net = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3))
    ref = torch.optim.Adam([p for p in net.parameters()], lr=5e-4, weight_decay=1e-5)
    for _ in range(3):
        ref.zero_grad()
        net(torch.randn(4, 7)).square().sum().backward()
        ref.step()
    sd = ref.state_dict()
    mine = OPT.FlatAdam(net.parameters(), lr=1.0)
    mine.load_state_dict(sd)
    assert mine.step_count == 3
    assert mine.param_groups[0]["lr"] == 5e-4 and mine.param_groups[0]["weight_decay"] == 1e-5
    back = mine.state_dict()
    assert back["param_groups"][0]["params"] == sd["param_groups"][0]["params"]
    assert set(back["state"]) == set(sd["state"])
    for i, st in sd["state"].items():
        assert torch.equal(back["state"][i]["exp_avg"], st["exp_avg"])
        assert torch.equal(back["state"][i]["exp_avg_sq"], st["exp_avg_sq"])
        assert float(back["state"][i]["step"]) == float(st["step"])
    # ... and torch accepts what FlatAdam wrote (a reference run can resume from our checkpoint)
    ref2 = torch.optim.Adam([p for p in net.parameters()], lr=1.0)
    ref2.load_state_dict(back)
    assert torch.equal(ref2.state_dict()["state"][0]["exp_avg"], sd["state"][0]["exp_avg"])
    # a fresh optimizer has no per-parameter state, like torch's
    assert OPT.FlatAdam(torch.nn.Linear(2, 2).parameters(), lr=1.0).state_dict()["state"] == {}
    with pytest.raises(ValueError):
        OPT.FlatAdam(torch.nn.Linear(2, 2).parameters(), lr=1.0).load_state_dict(sd)


@pytest.mark.parametrize("reward", ["gumbel", "reinforce"])
def test_joint_save_then_resume(tmp_path, reward):
    m = _Agents(share=False)
    opt = _opt(retrieval_reward=reward, checkpoint_path=str(tmp_path))
    d = OPT.load_optimizer(m, opt)
    if reward == "reinforce":
        assert set(d) == {"speaker", "listener"} and opt.alternating_turn == ["speaker", "listener"]
        spk, lis = d["speaker"], d["listener"]
    else:
        # both agents step in the speaker turn; the listener turn is gone (optimizer.py:88-95)
        assert set(d) == {"speaker"} and set(d["speaker"]) == {"speaker", "listener"}
        assert opt.alternating_turn == ["speaker"]
        spk, lis = d["speaker"]["speaker"], d["speaker"]["listener"]
    assert len(spk.params) == len(list(m.caption_generator.parameters()))
    assert len(lis.params) == len(list(m.vse.parameters()))
    _fake_moments(spk, 3)
    _fake_moments(lis, 4)
    OPT.save_optimizer(opt, d)
    assert sorted(os.listdir(tmp_path)) == ["listener_optimizer.pth", "speaker_optimizer.pth"]
    # the files hold torch.optim-shaped dictionaries
    sd = torch.load(tmp_path / "speaker_optimizer.pth")
    assert set(sd) == {"state", "param_groups"} and sd["param_groups"][0]["amsgrad"] is False
    m2 = _Agents(share=False)
    opt2 = _opt(retrieval_reward=reward, start_from=str(tmp_path))
    d2 = OPT.load_optimizer(m2, opt2)
    spk2, lis2 = (d2["speaker"], d2["listener"]) if reward == "reinforce" else \
        (d2["speaker"]["speaker"], d2["speaker"]["listener"])
    for a, b in ((spk, spk2), (lis, lis2)):
        assert b.step_count == 7 and _same_moments(a, b)


def test_joint_resume_falls_back_to_the_pretraining_optimizers(tmp_path):
    """No <turn>_optimizer.pth yet: the speaker resumes from --speaker_stage_2_optimizer_path and
    the listener from optimizer.pth next to --initialize_retrieval (optimizer.py:60-64,81-86)."""
    m = _Agents(share=False)
    pre = tmp_path / "pre"
    phase1 = tmp_path / "phase1"
    pre.mkdir(), phase1.mkdir()
    s = OPT.define_optimizer(m.caption_generator, _opt())
    _fake_moments(s, 5)
    torch.save(s.state_dict(), pre / "stage2.pth")
    li = OPT.define_optimizer(m.vse, _opt())
    _fake_moments(li, 6)
    torch.save(li.state_dict(), phase1 / "optimizer.pth")
    start = tmp_path / "joint"
    start.mkdir()
    m2 = _Agents(share=False)
    d = OPT.load_optimizer(m2, _opt(start_from=str(start),
                                    speaker_stage_2_optimizer_path=str(pre / "stage2.pth"),
                                    initialize_retrieval=str(phase1 / "model_vse-best.pth")))
    assert _same_moments(d["speaker"]["speaker"], s)
    assert _same_moments(d["speaker"]["listener"], li)


@pytest.mark.parametrize("phase,agent", [(1, "vse"), (2, "caption_generator"), (3, "caption_generator")])
def test_single_agent_phases(tmp_path, phase, agent):
    m = _Agents(share=False)
    opt = _opt(is_alternating=0, alternating_turn=None, phase=phase, checkpoint_path=str(tmp_path))
    d = OPT.load_optimizer(m, opt)
    assert set(d) == {"optimizer"}
    assert len(d["optimizer"].params) == len(list(getattr(m, agent).parameters()))
    _fake_moments(d["optimizer"], 9)
    OPT.save_optimizer(opt, d)
    assert os.listdir(tmp_path) == ["optimizer.pth"]
    d2 = OPT.load_optimizer(_Agents(share=False), _opt(is_alternating=0, alternating_turn=None,
                                                        phase=phase, start_from=str(tmp_path)))
    assert _same_moments(d2["optimizer"], d["optimizer"]) and d2["optimizer"].step_count == 7
    assert OPT.load_optimizer_path(_opt(is_alternating=0)) is None


def test_shared_embedding_is_aliased_not_orphaned():
    """with share_embed both agents' optimizers hold the embedding.  The second one must
    update the SAME storage with its own moments, not take the parameter over."""
    m = _Agents(share=True)
    emb = m.vse["embed"].weight
    assert m.caption_generator["embed"].weight is emb
    w0 = emb.detach().clone()
    spk = OPT.define_optimizer(m.caption_generator, _opt())
    lis = OPT.define_optimizer(m.vse, _opt())
    i_s = [i for i, p in enumerate(spk.params) if p is emb][0]
    i_l = [i for i, p in enumerate(lis.params) if p is emb][0]
    assert not spk.foreign[i_s] and lis.foreign[i_l]
    assert torch.equal(emb, w0)
    # the parameter lives in the speaker's bucket ...
    lo = spk.flat_param.data_ptr()
    assert lo <= emb.data_ptr() < lo + 4 * spk.flat_param.numel()
    # ... and the listener's own bucket holds only its own parameters, its moments cover both
    assert lis.flat_param.numel() == lis.own_numel < lis.numel == lis.exp_avg.numel()
    assert lis.offsets[i_l] >= lis.own_numel
    # both optimizers see the gradient of one backward pass
    for o in (spk, lis):
        o.zero_grad()
    (emb.sum() * 2.0).backward()
    spk.gather_grads()
    lis.gather_grads()
    for o, i in ((spk, i_s), (lis, i_l)):
        seg = o.flat_grad[o.offsets[i]:o.offsets[i] + emb.numel()]
        assert float(seg.min()) == 2.0 == float(seg.max())
    # checkpoints keep torch's layout: the shared parameter has an entry in both
    spk.step_count = lis.step_count = 1
    assert len(lis.state_dict()["state"]) == len(list(m.vse.parameters()))
