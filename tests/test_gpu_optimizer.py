"""GPU: coopcap_clamp_adam / FlatAdam against the reference's optimizer policy, restated with
torch itself: `g.clamp_(-c, c)` (misc/utils.py:65-69) then `torch.optim.Adam(lr, weight_decay)`
with default betas / eps (optimizer.py:25-27), applied by update_optimizer (optimizer.py:233-242).
fp32 elementwise arithmetic: parameters and both moments must agree to 1e-6 relative."""
import argparse

import pytest
import torch

pytestmark = pytest.mark.gpu

SPEAKER_PLUS_LISTENER = 26_133_777      # 14 452 497 + 11 681 280 parameters


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("n,weight_decay,world", [(SPEAKER_PLUS_LISTENER, 0.0, 1),
                                                  (SPEAKER_PLUS_LISTENER, 1e-4, 8),
                                                  (1_000_003, 1e-4, 8),      # odd length: scalar tail path
                                                  (3, 0.0, 1)])
def test_clamp_adam_matches_torch_for_five_steps(n, weight_decay, world):
    from cooperativeimagecaptioning_b200 import engine as EN
    g = torch.Generator(device="cuda").manual_seed(n % 1000 + world)
    clip, lr = 0.1, 5e-4
    p = torch.randn(n, device="cuda", generator=g) * 0.1
    ref_p = torch.nn.Parameter(p.clone())
    ref = torch.optim.Adam([ref_p], lr=lr, weight_decay=weight_decay)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 6):
        # gradient magnitudes straddle the clamp; with `world` ranks the bucket holds the SUM
        grad_mean = torch.randn(n, device="cuda", generator=g) * (0.02 * step)
        summed = grad_mean * world
        EN.clamp_adam_(p, summed, m, v, step=step, lr=lr, grad_scale=1.0 / world, clip=clip,
                       weight_decay=weight_decay)
        ref_p.grad = (summed / world).clamp_(-clip, clip)
        ref.step()
        st = ref.state[ref_p]
        assert _rel(m, st["exp_avg"]) <= 1e-6, step
        assert _rel(v, st["exp_avg_sq"]) <= 1e-6, step
        assert _rel(p, ref_p.data) <= 1e-6, step
        assert float((p - ref_p.data).abs().max()) <= 2e-7, step
    assert int((summed.abs() / world > clip).sum()) > 0 or n < 100     # the clamp was exercised


class _Agents(torch.nn.Module):
    def __init__(self, share):
        super().__init__()
        torch.manual_seed(0)
        self.vse = torch.nn.ModuleDict(dict(embed=torch.nn.Embedding(37, 16), fc=torch.nn.Linear(16, 9)))
        self.caption_generator = torch.nn.ModuleDict(
            dict(embed=torch.nn.Embedding(37, 16), out=torch.nn.Linear(16, 5)))
        if share:
            self.caption_generator["embed"] = self.vse["embed"]

    def forward(self, ids):
        e = self.caption_generator["embed"](ids)
        f = self.vse["embed"](ids)
        return self.caption_generator["out"](e).square().sum() + 3.0 * self.vse["fc"](f).sum()


@pytest.mark.parametrize("reward", ["gumbel", "reinforce"])
def test_shared_embedding_steps_like_two_torch_adams(reward):
    """ADVICE r1 (--share_embed 1): the embedding is in both agents' optimizers; the reference
    applies both Adam updates in the gumbel / multinomial turns (optimizer.py:233-237) and the
    active agent's in a REINFORCE turn.  Same through load_optimizer / update_optimizer here."""
    from cooperativeimagecaptioning_b200 import optimizer as OPT
    opt = argparse.Namespace(is_alternating=1, alternating_turn=["speaker", "listener"],
                             retrieval_reward=reward, start_from=None, share_embed=1,
                             learning_rate=5e-3, weight_decay=1e-4, grad_clip=0.1, phase=None)
    mine, ref = _Agents(True).cuda(), _Agents(True).cuda()
    ref.load_state_dict(mine.state_dict())
    od = OPT.load_optimizer(mine, opt)
    r_spk = torch.optim.Adam(list(ref.caption_generator.parameters()), lr=5e-3, weight_decay=1e-4)
    r_lis = torch.optim.Adam(list(ref.vse.parameters()), lr=5e-3, weight_decay=1e-4)
    gen = torch.Generator().manual_seed(3)
    turns = ["speaker"] * 4 if reward != "reinforce" else ["speaker", "listener", "speaker", "listener"]
    for turn in turns:
        ids = torch.randint(0, 37, (12,), generator=gen).cuda()
        optimizer = od[turn]
        OPT.zeroing_optimizer(opt, od, optimizer)
        mine(ids).backward()
        OPT.update_optimizer(od, optimizer, opt)
        active = [r_spk, r_lis] if reward != "reinforce" else [r_spk if turn == "speaker" else r_lis]
        for o in active:
            o.zero_grad()
        ref(ids).backward()
        for o in active:            # clip_gradient is idempotent, so clamping twice == once
            for p in o.param_groups[0]["params"]:
                if p.grad is not None:
                    p.grad.clamp_(-0.1, 0.1)
            o.step()
        for (n, a), (_, b) in zip(mine.named_parameters(), ref.named_parameters()):
            assert _rel(a.data, b.data) <= 2e-6, (turn, n)
    assert mine.caption_generator["embed"].weight is mine.vse["embed"].weight
    # the shared embedding really moved (it was orphaned before the fix)
    fresh = _Agents(True).cuda()
    assert float((mine.vse["embed"].weight - fresh.vse["embed"].weight).abs().max()) > 1e-3


def test_learning_rate_schedule_of_train_py_drives_the_optimizers():
    """train.py:50-76 `update_learning_rate`: the decayed rate is written into
    `optimizer.param_groups[*]['lr']` of every optimizer of the nested dict (misc/utils.py:60-62
    `set_lr`) between steps.  FlatAdam must take its step size from there, like torch's Adam."""
    from cooperativeimagecaptioning_b200 import optimizer as OPT
    opt = argparse.Namespace(is_alternating=1, alternating_turn=["speaker", "listener"],
                             retrieval_reward="gumbel", start_from=None, share_embed=0,
                             learning_rate=5e-3, weight_decay=0.0, grad_clip=0.1, phase=None,
                             learning_rate_decay_start=0, learning_rate_decay_every=1,
                             learning_rate_decay_rate=0.5)
    mine, ref = _Agents(False).cuda(), _Agents(False).cuda()
    ref.load_state_dict(mine.state_dict())
    od = OPT.load_optimizer(mine, opt)
    r_spk = torch.optim.Adam(list(ref.caption_generator.parameters()), lr=5e-3)
    r_lis = torch.optim.Adam(list(ref.vse.parameters()), lr=5e-3)
    gen = torch.Generator().manual_seed(5)
    for epoch in range(1, 5):
        # the reference's loop body, verbatim in structure (gumbel case: every optimizer under 'speaker')
        frac = (epoch - opt.learning_rate_decay_start) // opt.learning_rate_decay_every
        opt.current_lr = opt.learning_rate * opt.learning_rate_decay_rate ** frac
        for agent_in in od["speaker"].keys():
            OPT.set_lr(od["speaker"][agent_in], opt.current_lr)
        for o in (r_spk, r_lis):
            for group in o.param_groups:
                group["lr"] = opt.current_lr
        ids = torch.randint(0, 37, (12,), generator=gen).cuda()
        optimizer = od["speaker"]
        OPT.zeroing_optimizer(opt, od, optimizer)
        mine(ids).backward()
        OPT.update_optimizer(od, optimizer, opt)
        for o in (r_spk, r_lis):
            o.zero_grad()
        ref(ids).backward()
        for o in (r_spk, r_lis):
            for p in o.param_groups[0]["params"]:
                if p.grad is not None:
                    p.grad.clamp_(-0.1, 0.1)
            o.step()
        for (n, a), (_, b) in zip(mine.named_parameters(), ref.named_parameters()):
            assert _rel(a.data, b.data) <= 2e-6, (epoch, n)
    assert abs(opt.current_lr - 5e-3 * 0.5 ** 4) < 1e-12
