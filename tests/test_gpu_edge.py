"""GPU: decode bit-exactness, listener variants, shared embedding and edge shapes -- the cases the
reference's call sites exercise beyond the main joint step."""
import pytest
import torch

from oracle import joint as OJ
from oracle import listener as OL
from oracle import speaker as OS
from oracle import synth
from oracle.ref_loader import reference_opt
from gpu_util import REAL, cuda_params

pytestmark = pytest.mark.gpu


def _models(B, seed, eos_bias=0.0, **optkw):
    import cooperativeimagecaptioning_b200.models as models
    d = REAL
    Ps = synth.speaker_params(d, seed=seed, eos_bias=eos_bias)
    Pl = synth.listener_params(d, seed=seed + 1)
    opt = reference_opt(batch_size=B, **optkw)
    m = models.AlternatingJointModel(opt)
    sd = {"caption_generator." + k: v for k, v in Ps.items()}
    sd.update({"vse." + k: v for k, v in Pl.items()})
    m.load_state_dict(sd)
    return m.cuda(), Ps, Pl


@pytest.mark.parametrize("B,L,varlen", [(16, 36, False), (5, 3, True), (1, 1, False), (512, 36, False)])
def test_greedy_decode_is_bit_exact_off_near_ties(B, L, varlen):
    """Config 3: eval-mode greedy captions (AttModel.py:327-329).  A row may diverge from the
    oracle only at a step where the oracle's top-2 log-prob gap is below the near-tie tolerance;
    rows are compared up to their first such step."""
    d = REAL
    m, Ps, _ = _models(B, 81, eos_bias=0.0)
    m.eval()
    batch = synth.make_batch(d, B, L, 83, varlen=varlen, min_regions=1)
    ref = OS.sample(Ps, batch.att_feats, batch.att_masks, mode="reinforce", seq_length=d.seq_length,
                    vocab_size=d.vocab_size, noise=OS.SpeakerNoise(), drop_p=0.0, sample_max=1,
                    keep_all_steps=True)
    with torch.no_grad():
        seq, logp = m.sample(batch.fc_feats.cuda(), batch.att_feats.cuda(),
                             None if batch.att_masks is None else batch.att_masks.cuda(),
                             {"sample_max": 1})
    seq, logp = seq.cpu(), logp.cpu()
    ref_raw = torch.stack(ref.tokens_raw, 1)                           # [B, 16] before masking
    n = seq.shape[1]
    exact, near, worst_gap = 0, 0, 0.0
    for b in range(B):
        alive = True
        for t in range(n):
            if not alive:
                break
            want = int(ref_raw[b, t])
            if int(seq[b, t]) != want:
                top2 = ref.step_logprobs[t][b].topk(2)[0]
                assert float(top2[0] - top2[1]) < 5e-3, (b, t, float(top2[0] - top2[1]))
                worst_gap = max(worst_gap, float(top2[0] - top2[1]))
                near += 1
                break
            exact += 1
            ref_lp = float(ref.step_logprobs[t][b, want])
            assert abs(float(logp[b, t]) - ref_lp) <= 2e-2 * abs(ref_lp)
            alive = want > 0
    print(f"greedy decode B={B} L={L}: {exact} tokens bit-exact, {near} rows stopped at a near-tie")
    assert exact >= B          # at least the first token of every row
    if B == 512:               # BASELINE.json configs[2] at its size: keep the record
        from gpu_util import write_report
        write_report("config3_greedy_decode_512x36", dict(
            case="config3_greedy_decode_512x36", rows=B, regions=L, tokens_bit_exact=exact,
            rows_stopped_at_a_near_tie=near, largest_gap_at_a_stop=worst_gap, near_tie_gap_tolerance=5e-3,
            logprob_rel_tolerance=2e-2,
            note="free-running greedy decode with random-initialised weights: the 9488 logits of a row are "
                 "nearly flat, so top-2 gaps below the bf16 resolution are frequent; every divergence is "
                 "checked to sit at such a gap, rows are compared up to their first one"))


@pytest.mark.parametrize("pool,use_abs,max_violation,whole_batch,only", [
    ("mean", 1, 1, False, "off"),
    ("max", 0, 0, True, "off"),
    ("last", 0, 0, False, "image"),
    ("max", 1, 1, True, "caption"),
    ("mean", 0, 0, False, "off"),
])
def test_listener_option_variants(pool, use_abs, max_violation, whole_batch, only):
    """vse_pool_type mean / max (VSEFCModel.py:115-126), vse_use_abs (:50-52,:137-139) and
    vse_max_violation = 0 (mean over the negatives, :190-193): loss and parameter gradients.
    The options' non-smooth decisions (sign of |.|, max-pool arg-max step, hardest negative) are
    replayed from the CUDA pass and may differ from the oracle's own only at near-ties."""
    d = REAL
    B = 9
    m, _, Pl = _models(B, 95, vse_pool_type=pool, vse_use_abs=use_abs, vse_max_violation=max_violation)
    m.vse.keep_passes = True
    batch = synth.make_batch(d, B, 4, 97)
    out = m.vse(batch.fc_feats.cuda(), None, batch.labels.cuda(), batch.masks.cuda(), whole_batch,
                only_one_retrieval=only)
    lp = m.vse._passes[0]
    enc, hr = {}, None
    if use_abs:
        enc["im_sign"] = torch.where(lp.t["img_pre"].cpu() < 0, -1.0, 1.0)
        pre = lp.t["cap_pre"] if pool != "last" else lp.t["h32"][-1]
        enc["cap_sign"] = torch.where(pre.cpu() < 0, -1.0, 1.0)
    if pool == "max":
        enc["pool_arg"] = lp.t["pool_arg"].cpu().long()
    if max_violation:
        hr = {"arg_s": lp.t["arg_s"].cpu().long(), "arg_im": lp.t["arg_im"].cpu().long()}
    Plo = {k: v.clone().requires_grad_(True) for k, v in Pl.items()}
    ref = OL.vse_forward(Plo, batch.fc_feats, batch.labels, batch.masks, whole_batch, only, 0.2,
                         bool(max_violation), pool, use_abs=bool(use_abs), hinge_replay=hr,
                         enc_replay=enc)
    # replayed decisions differ from the oracle's own only at near-ties
    flips = {}
    for key in ("im", "cap"):
        if key + "_sign" in enc:
            v = enc[key + "_val"]
            bad = (v < 0) != (enc[key + "_sign"] < 0)
            flips[key] = int(bad.sum())
            assert not bool(bad.any()) or float(v[bad].abs().max()) <= 5e-2 * float(v.abs().median())
    if pool == "max":
        vals = enc["pool_vals"]
        own = vals.max(1)[0]
        got = vals.gather(1, enc["pool_arg"][:, None, :]).squeeze(1)
        flips["pool"] = int((own != got).sum())
        assert float((own - got).abs().max()) <= 5e-2 * float(own.abs().median())
    if hr is not None:
        from gpu_util import check_hinge_near_ties
        flips.update(check_hinge_near_ties(hr))
    print(f"listener variant {pool}/{use_abs}/{max_violation}: decision flips {flips}")
    w = torch.linspace(0.5, 1.5, B)
    ref_scalar = (ref * w).sum() if whole_batch else ref
    gref = torch.autograd.grad(ref_scalar, list(Plo.values()), allow_unused=True)
    assert out.shape == ref.shape
    assert torch.allclose(out.detach().cpu(), ref.detach(), rtol=2e-2, atol=2e-3), (out, ref)
    ((out * w.cuda()).sum() if whole_batch else out).backward()
    for (k, _), g in zip(Plo.items(), gref):
        p = dict(m.vse.named_parameters())[k]
        got = torch.zeros(p.numel(), dtype=torch.double) if p.grad is None else \
            p.grad.cpu().double().flatten()
        r = torch.zeros_like(got) if g is None else g.double().flatten()
        if float(r.norm()) == 0:
            assert float(got.norm()) == 0, k
            continue
        # 9 rows: the bias gradients are column sums of bf16-rounded per-row terms that largely
        # cancel (most of all under the dense sum-violation hinge)
        tol = 8e-2 if k.endswith("bias") or "bias_" in k else 3e-2
        err = float((got - r).norm() / r.norm())
        assert err <= tol, (k, err)


def test_decoding_constraint_bans_the_previous_word():
    """AttModel.sample with opt['decoding_constraint'] = 1 (AttModel.py:437-442): the logit of the
    previously emitted id is -inf.  Greedy ids bit-exact against the oracle off near-ties, no
    immediate repeats, and the same weights without the constraint do repeat (so it is active)."""
    d = REAL
    B, L = 12, 5
    m, Ps, _ = _models(B, 85, eos_bias=-8.0)
    m.eval()
    batch = synth.make_batch(d, B, L, 87, varlen=True, min_regions=1)
    ref = OS.sample(Ps, batch.att_feats, batch.att_masks, mode="reinforce", seq_length=d.seq_length,
                    vocab_size=d.vocab_size, noise=OS.SpeakerNoise(), drop_p=0.0, sample_max=1,
                    keep_all_steps=True, decoding_constraint=1)
    fc, att, am = batch.fc_feats.cuda(), batch.att_feats.cuda(), batch.att_masks.cuda()
    with torch.no_grad():
        seq, logp = m.sample(fc, att, am, {"sample_max": 1, "decoding_constraint": 1})
        free, _ = m.sample(fc, att, am, {"sample_max": 1})
    seq, logp, free = seq.cpu(), logp.cpu(), free.cpu()
    rep = lambda s: int(((s[:, 1:] == s[:, :-1]) & (s[:, 1:] > 0)).sum())
    assert rep(seq) == 0
    assert rep(free) > 0, "the unconstrained decode has no repeats: the case does not exercise the ban"
    ref_raw = torch.stack(ref.tokens_raw, 1)
    exact = 0
    for b in range(B):
        for t in range(seq.shape[1]):
            want = int(ref_raw[b, t])
            if int(seq[b, t]) != want:
                top2 = ref.step_logprobs[t][b].topk(2)[0]
                assert float(top2[0] - top2[1]) < 5e-3, (b, t, float(top2[0] - top2[1]))
                break
            exact += 1
            ref_lp = float(ref.step_logprobs[t][b, want])
            assert abs(float(logp[b, t]) - ref_lp) <= 2e-2 * abs(ref_lp)
            if want == 0:
                break
    print(f"constrained greedy decode: {exact} tokens bit-exact, {rep(free)} repeats without the ban")
    assert exact >= 4 * B


@pytest.mark.parametrize("only,whole_batch", [("off", False), ("image", False), ("caption", True),
                                              ("off", True)])
def test_listener_variants(only, whole_batch):
    """VSEFCModel.forward with only_one_retrieval / whole_batch (VSEFCModel.py:197-207), loss and
    parameter gradients."""
    d = REAL
    B = 9
    m, _, Pl = _models(B, 91)
    m.vse.keep_passes = True
    batch = synth.make_batch(d, B, 4, 93)
    Plo = {k: v.clone().requires_grad_(True) for k, v in Pl.items()}
    ref = OL.vse_forward(Plo, batch.fc_feats, batch.labels, batch.masks, whole_batch, only)
    w = torch.linspace(0.5, 1.5, B)
    ref_scalar = (ref * w).sum() if whole_batch else ref
    gref = torch.autograd.grad(ref_scalar, list(Plo.values()))
    out = m.vse(batch.fc_feats.cuda(), None, batch.labels.cuda(), batch.masks.cuda(), whole_batch,
                only_one_retrieval=only)
    assert out.shape == ref.shape
    assert torch.allclose(out.detach().cpu(), ref.detach(), rtol=2e-2, atol=2e-3)
    # the max-violation arg-max (VSEFCModel.py:191-193) is a non-smooth decision: it must agree
    # with the oracle's, or the oracle's top-2 violations must be a near-tie
    with torch.no_grad():
        S = OL.img_enc(Pl, batch.fc_feats) @ OL.txt_enc(Pl, batch.labels, batch.masks).t()
        S = S.masked_fill(torch.eye(B, dtype=torch.bool), -1e9)
    lp = m.vse._passes[0]
    for dim, key in ((1, "arg_s"), (0, "arg_im")):
        got_arg = lp.t[key].cpu().long()
        top2 = S.topk(2, dim=dim)
        ref_arg = top2[1].select(dim, 0)
        gap = (top2[0].select(dim, 0) - top2[0].select(dim, 1))
        bad = got_arg != ref_arg
        assert bool((gap[bad] < 2e-3).all()), (key, gap[bad])
        assert not bool(bad.any()), "near-tie flip in this fixed case: pick another seed"
    ((out * w.cuda()).sum() if whole_batch else out).backward()
    for (k, _), g in zip(Plo.items(), gref):
        got = dict(m.vse.named_parameters())[k].grad.cpu().double().flatten()
        r = g.double().flatten()
        if float(r.norm()) == 0:
            assert float(got.norm()) == 0
            continue
        # 9 rows only: less averaging of the bf16 rounding than in the 1024-row step
        assert float((got - r).norm() / r.norm()) <= 3e-2, k


def test_shared_embedding_accumulates_both_paths():
    """--share_embed (AlternatingJointModel.py:85-88): the speaker's input embedding IS the
    listener's embedding parameter; its gradient is the sum of both uses."""
    from cooperativeimagecaptioning_b200 import engine as EN
    from gpu_util import branch_replay, pack_keep, u8
    d = REAL
    B, L, T = 8, 5, d.seq_length
    m, Ps, Pl = _models(B, 101, eos_bias=7.5, share_embed=1, retrieval_reward="gumbel",
                        drop_prob_lm=0.0)
    m.train()
    spk = m.caption_generator
    assert spk.embed[0].weight is m.vse.txt_enc.embed.weight
    batch = synth.make_batch(d, B, L, 103)
    noise = synth.make_noise(d, B, L, 105, dropout=False, gumbel=True)
    shared = Pl["txt_enc.embed.weight"]
    Ps = dict(Ps)
    Ps["embed.0.weight"] = shared
    free = OS.sample(Ps, batch.att_feats, None, mode="gumbel", seq_length=T, vocab_size=d.vocab_size,
                     noise=noise, drop_p=0.0, sample_max=0, use_one_hot=1, keep_all_steps=True)
    forced = torch.stack(free.tokens_raw, 1)
    spk.injected = EN.SpeakerRandom(seed=1, drop_p=0.0, noise=noise.U.cuda().contiguous())
    spk.forced_tokens = forced.cuda()
    spk.keep_passes = True
    loss = m(batch.fc_feats.cuda(), batch.labels.cuda(), batch.masks.cuda(), None,
             batch.att_feats.cuda(), None, is_alternating=True, alternating_turn="speaker")
    loss.backward()
    rn = branch_replay(spk._passes[0], None, noise)
    sh = shared.clone().requires_grad_(True)
    Pso = {k: (sh if k == "embed.0.weight" else v.clone().requires_grad_(True)) for k, v in Ps.items()}
    Plo = {k: (sh if k == "txt_enc.embed.weight" else v.clone().requires_grad_(True)) for k, v in Pl.items()}
    cfg = OJ.JointCfg(drop_p=0.0, retrieval_reward="gumbel")
    loss_ref, *_ = OJ.st_joint_loss(Pso, Plo, batch.fc_feats, batch.att_feats, None, rn, cfg, forced)
    g = torch.autograd.grad(loss_ref, [sh])[0].double().flatten()
    got = m.vse.txt_enc.embed.weight.grad.cpu().double().flatten()
    assert abs(float(loss) - float(loss_ref)) <= 2e-2 * abs(float(loss_ref))
    assert float((got - g).norm() / g.norm()) <= 2e-2


def test_all_rows_finish_immediately():
    """Every row emits EOS at the first step: the reference crashes on torch.cat([])
    (Appendix D); here `sample` returns width-0 tensors and the joint loss stays finite."""
    B = 6
    m, _, _ = _models(B, 111, eos_bias=1e4, retrieval_reward="gumbel")
    m.train()
    d = REAL
    batch = synth.make_batch(d, B, 4, 113)
    fc, att = batch.fc_feats.cuda(), batch.att_feats.cuda()
    with torch.no_grad():
        seq, lp = m.sample(fc, att, None, {"sample_max": 1})
    assert seq.shape == (B, 0) and lp.shape == (B, 0)
    loss = m(fc, batch.labels.cuda(), batch.masks.cuda(), None, att, None, is_alternating=True,
             alternating_turn="speaker")
    loss.backward()
    assert torch.isfinite(loss)
    # captions are [BOS] only: no gradient reaches the speaker, the listener still trains
    assert all(p.grad is None or float(p.grad.abs().max()) == 0
               for n, p in m.caption_generator.named_parameters())
    assert float(m.vse.img_enc.fc.weight.grad.abs().max()) > 0
