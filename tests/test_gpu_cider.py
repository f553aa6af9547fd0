"""GPU: the CIDEr-D self-critical reward on the device (csrc/cider.cu through rewards.py) against
(a) golden vectors produced by the reference's own scorer (tests/golden/cider_*.npz), (b) the
float64 oracle restatement at the benchmark batch size, and (c) the joint step with the CIDEr term
(traditional_cider, AlternatingJointModel.py:407-431) against the oracle's loss and gradients."""
import numpy as np
import pytest
import torch

from oracle import cases
from oracle import cider as OC
from oracle import joint as OJ
from oracle import synth
from gpu_util import REAL, branch_replay, check_near_ties

pytestmark = pytest.mark.gpu

SCORE_TOL = 1e-9     # float64 scores: the sums run in the reference's order, only log / pow round differently
NEAR_TIE = 2e-2


def _tm(x):
    return torch.from_numpy(np.ascontiguousarray(x.T)).cuda()


def _scorer(df, ref_len):
    from cooperativeimagecaptioning_b200 import rewards
    return rewards.DeviceCiderD("corpus") if df is None else rewards.DeviceCiderD.from_table(df, ref_len)


@pytest.mark.parametrize("name", cases.cider_golden_names())
def test_device_scorer_matches_reference_golden(name):
    from cooperativeimagecaptioning_b200 import rewards
    meta, gts, gen, greedy, df, ref_len, z = cases.load_cider_golden(name)
    B = gen.shape[0]
    st = rewards.stage_gts(gts, B, torch.device("cuda"))
    res = rewards.reward_on_device(_scorer(df, ref_len), st, _tm(gen), _tm(greedy))
    s = res.scores.cpu().numpy()
    assert np.max(np.abs(s[:B] - z["out.cider_gen"])) <= SCORE_TOL
    assert np.max(np.abs((s[:B] - s[B:]) - z["out.reward"])) <= SCORE_TOL
    assert abs(float(s[B:].mean()) - float(z["out.cider_greedy"])) <= SCORE_TOL
    assert np.array_equal(res.reward.cpu().numpy(), z["out.reward"].astype(np.float32)) or \
        np.max(np.abs(res.reward.cpu().numpy() - z["out.reward"].astype(np.float32))) <= 1e-7
    # REINFORCE coefficients of traditional_cider; the caption width n is the longest sampled caption
    n = max(1, int((np.cumprod(gen > 0, 1)).sum(1).max()))
    want = OC.cider_loss_coef(z["out.reward"], gen[:, :n])
    got = res.coef.cpu().numpy().T
    assert np.max(np.abs(got[:, :n] - want)) <= 1e-6 * max(1.0, np.abs(want).max())
    assert not got[:, n:].any()
    st_ = res.stats.cpu().numpy()
    assert abs(st_[0] - z["out.reward"].mean()) <= 1e-9 and abs(st_[1] - float(z["out.cider_greedy"])) <= 1e-9
    # use_gen_cider_scores != 0: the sampled score itself is the reward (:415-419)
    res2 = rewards.reward_on_device(_scorer(df, ref_len), st, _tm(gen), _tm(greedy), differenced=False)
    assert np.max(np.abs(res2.reward.cpu().numpy() - z["out.cider_gen"].astype(np.float32))) <= 1e-7


def test_host_api_matches_reference_golden():
    """misc/rewards.py's own entry point (ids in, numpy out), as train-time callers use it."""
    from cooperativeimagecaptioning_b200 import rewards
    meta, gts, gen, greedy, df, ref_len, z = cases.load_cider_golden("cider_corpus_v30")
    rewards.CiderD_scorer = None
    rewards.init_scorer("corpus")
    try:
        n = max(1, int((np.cumprod(gen > 0, 1)).sum(1).max()))
        g = torch.from_numpy(gen[:, :n]).cuda()
        scores, greedy_mean = rewards.get_self_critical_reward({"gts": gts}, g, torch.from_numpy(greedy).cuda())
        assert np.max(np.abs(scores - z["out.reward"])) <= SCORE_TOL
        assert abs(greedy_mean - float(z["out.cider_greedy"])) <= SCORE_TOL
        cg, sc, gm = rewards.get_self_critical_reward({"gts": gts}, g, torch.from_numpy(greedy).cuda(),
                                                      return_gen_scores=True)
        assert np.max(np.abs(cg - z["out.cider_gen"])) <= SCORE_TOL
        assert rewards.array_to_str([3, 5, 0, 7]) == "3 5 0"
    finally:
        rewards.CiderD_scorer = None


def test_device_scorer_at_benchmark_size():
    """1020 rows = 204 images x 5, vocabulary 9487, against the oracle (a few seconds on CPU)."""
    from cooperativeimagecaptioning_b200 import rewards
    rng = np.random.default_rng(7)
    B, spi, V = 1020, 5, 9487                      # 204 images x 5 rows
    gts, gen, greedy = [], np.zeros((B, 16), np.int64), np.zeros((B, 16), np.int64)
    for i in range(B // spi):
        caps = np.zeros((int(rng.integers(5, 8)), 16), np.int64)
        for c in caps:
            k = int(rng.integers(5, 17))
            c[:k] = rng.integers(1, 60 if rng.random() < 0.5 else V + 1, size=k)
        gts.append(caps)
    for b in range(B):
        g = gts[b // spi]
        for arr in (gen, greedy):
            if rng.random() < 0.6:
                arr[b] = g[int(rng.integers(0, len(g)))]
                for _ in range(int(rng.integers(0, 4))):
                    arr[b, int(rng.integers(0, 16))] = int(rng.integers(0, 60))
            else:
                k = int(rng.integers(1, 17))
                arr[b, :k] = rng.integers(1, 60, size=k)
        # finished rows hold zeros after their first 0, as the decode loop leaves them
        for arr in (gen, greedy):
            z = np.where(arr[b] == 0)[0]
            if len(z):
                arr[b, z[0]:] = 0
    st = rewards.stage_gts(gts, B, torch.device("cuda"))
    res = rewards.reward_on_device(rewards.DeviceCiderD("corpus"), st, _tm(gen), _tm(greedy))
    cg, diff, gm = OC.self_critical_reward(gts, gen, greedy)
    s = res.scores.cpu().numpy()
    assert np.count_nonzero(cg) > B // 2
    assert np.max(np.abs(s[:B] - cg)) <= SCORE_TOL and np.max(np.abs((s[:B] - s[B:]) - diff)) <= SCORE_TOL


def _gts_from_forced(forced, forced_g, spi, seed, V):
    """Ground-truth sets that overlap the captions the test will sample: image i gets the sampled
    caption of one of its rows (perturbed), the greedy caption of another, and a random one."""
    rng = np.random.default_rng(seed)
    B, T = forced.shape
    cut = lambda r: np.where(np.cumprod(r > 0) > 0, r, 0)
    gts = []
    for i in range(B // spi):
        a = cut(forced[i * spi].numpy().copy())
        a[int(rng.integers(0, 3))] = int(rng.integers(1, V + 1))
        b = cut(forced_g[i * spi + spi - 1].numpy().copy())
        c = np.zeros(T, np.int64)
        k = int(rng.integers(4, T))
        c[:k] = rng.integers(1, V + 1, size=k)
        d = cut(forced[i * spi + 1].numpy().copy()) if spi > 1 else c
        gts.append(np.stack([a, b, c, d], 0))
    return gts


@pytest.mark.parametrize("mode,weight,use_gen", [("gumbel", 0.01, 0), ("reinforce", 0.8, 0), ("reinforce", 0.0, 1)])
def test_joint_step_with_cider_term(mode, weight, use_gen):
    """Speaker turn with cider_optimization > 0: Gumbel joint step + CIDEr term on the same pass,
    REINFORCE (gt baseline) + CIDEr, and the CIDEr term alone (gen_result_for_cider)."""
    from test_gpu_models import _build, _check_grads, _oracle_grads, _replay_tokens, LOSS_TOL
    from oracle import speaker as OS
    B, L, spi, cider_w = 12, 8, 3, 0.5
    model, Ps, Pl, batch, noise, cfg = _build(
        mode, B, L, 61, varlen=(mode == "gumbel"), dropout=False, eos_bias=3.0,
        retrieval_reward_weight=weight, cider_optimization=cider_w, use_gen_cider_scores=use_gen,
        reinforce_baseline_type="gt", is_alternating=1)
    d = REAL
    forced = _replay_tokens(Ps, batch, noise, mode, 0.0, 1.0)
    forced_g = _replay_tokens(Ps, batch, synth.SpeakerNoise(), "reinforce", 0.0, 1.0, sample_max=1)
    gts = _gts_from_forced(forced, forced_g, spi, 5, d.vocab_size)
    spk = model.caption_generator
    orig = spk._sample_pass

    def patched(att_feats, att_masks, sample_max, temperature, use_one_hot, **kw):
        spk.forced_tokens = (forced_g if sample_max else forced).cuda()
        return orig(att_feats, att_masks, sample_max, temperature, use_one_hot, **kw)
    spk._sample_pass = patched
    loss = model(batch.fc_feats.cuda(), batch.labels.cuda(), batch.masks.cuda(), {"gts": gts},
                 batch.att_feats.cuda(), None if batch.att_masks is None else batch.att_masks.cuda(),
                 is_alternating=True, alternating_turn="speaker")
    loss.backward()
    rn = branch_replay(spk._passes[0], batch.att_masks, noise)
    Pso = {k: v.clone().requires_grad_(True) for k, v in Ps.items()}
    Plo = {k: v.clone().requires_grad_(True) for k, v in Pl.items()}
    if mode == "gumbel":
        loss_ref, res, _, _ = OJ.st_joint_loss(Pso, Plo, batch.fc_feats, batch.att_feats, batch.att_masks,
                                               rn, cfg, forced)
    elif weight > 0:
        cfg.vse_loss_weight = 0.0
        loss_ref, res, _, _ = OJ.reinforce_speaker_loss(
            Pso, Plo, batch.fc_feats, batch.att_feats, batch.att_masks, batch.labels, batch.masks, rn,
            cfg, forced_tokens=forced)
    else:
        res = OS.sample(Pso, batch.att_feats, batch.att_masks, mode="reinforce", seq_length=d.seq_length,
                        vocab_size=d.vocab_size, noise=rn, drop_p=0.0, sample_max=0, temperature=1.0,
                        forced_tokens=forced)
        loss_ref = 0.0
    g = OJ.greedy_for_cider(Ps, batch.att_feats, batch.att_masks, synth.SpeakerNoise(), cfg, forced_g)
    loss_cider, reward, cider_greedy = OJ.cider_term(res.logprobs, res.seq, g.seq, gts,
                                                     use_gen_cider_scores=use_gen)
    loss_ref = loss_ref + cider_w * loss_cider
    ref = _oracle_grads(loss_ref, Pso, Plo)
    print(check_near_ties(rn, batch.att_masks, NEAR_TIE))
    got_r = model._cider_last.reward.cpu().numpy()
    assert np.count_nonzero(reward) >= B // 3, "degenerate test: rewards are all zero"
    assert np.max(np.abs(got_r - reward.astype(np.float32))) <= 1e-6
    out = model.loss()
    assert abs(float(out["loss_cider"]) - float(loss_cider)) <= LOSS_TOL * max(abs(float(loss_cider)), 1e-3)
    assert abs(float(out["avg_reward"]) - reward.mean()) <= 1e-6
    assert abs(float(out["cider_greedy"]) - cider_greedy) <= 1e-6
    denom = max(abs(float(loss_ref)), 1e-3)
    assert abs(float(loss) - float(loss_ref)) <= LOSS_TOL * denom, (float(loss), float(loss_ref))
    _check_grads(model, ref, f"cider-{mode}-{weight}")
