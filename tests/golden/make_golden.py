"""Generate the golden vectors in this directory by running the REAL reference.

Runs only in the development container (needs /root/reference):

    python tests/golden/make_golden.py

For every case it (1) builds seeded synthetic parameters / inputs / noise (oracle/synth.py),
(2) drives the reference's own AlternatingJointModel / AttModel / VSEFCModel -- imported in memory
through oracle/ref_loader.py -- with that noise injected at the points SURVEY.md §8(c) lists
(sample_gumbel, torch.multinomial, F.dropout, the partial-sampling uniforms), (3) runs the oracle
restatement on the same tensors and asserts agreement (this is what pins the oracle), and
(4) stores inputs + reference outputs as tests/golden/<case>.npz.  tests/test_oracle_golden.py
re-checks the oracle against these files on every CPU run, with no reference tree needed.
"""
from __future__ import annotations

import json
import os
import sys
import warnings
from dataclasses import asdict

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import joint as OJ  # noqa: E402
from oracle import ref_loader  # noqa: E402
from oracle import speaker as OS  # noqa: E402
from oracle import synth  # noqa: E402

warnings.filterwarnings("ignore")


class Inject:
    """Feeds pre-drawn noise into the reference's random call sites."""

    def __init__(self, ref_models, noises, rows, att_masks):
        # `from .AttModel import *` shadows the sampler submodules with functions of the same
        # name on the package, so fetch the modules from sys.modules
        self.m_gumbel = sys.modules["models.gumbel"]
        self.m_gumbel_soft = sys.modules["models.gumbel_softmax"]
        self.noises = list(noises)
        self.rows = rows
        self.att_masks = att_masks
        self.cur = -1
        self.step_calls = 0
        self.t_gumbel = 0
        self.t_mult = 0
        self.t_part = 0

    # -- F.dropout ---------------------------------------------------------------------------
    def _packed(self, keep):
        lens = self.att_masks.long().sum(1)
        sorted_lens, idx = torch.sort(lens, descending=True)          # AttModel.py:32
        chunks = []
        for t in range(int(sorted_lens[0])):
            bs = int((sorted_lens > t).sum())
            chunks.append(keep[idx[:bs], t])
        return torch.cat(chunks, 0)

    def dropout(self, x, p=0.5, training=True, inplace=False):
        if not training or p == 0.0:
            return x
        is_step = x.dim() == 2 and x.size(0) == self.rows
        if not is_step:                      # att_embed: first dropout call of a speaker pass
            self.cur += 1
            self.step_calls = 0
            self.t_gumbel = self.t_mult = self.t_part = 0
            keep = self.noises[self.cur].drop_att
            if x.dim() == 2:                 # PackedSequence.data order
                keep = self._packed(keep)
            else:
                keep = keep[:, : x.size(1)]
        else:
            n = self.noises[self.cur]
            t, which = divmod(self.step_calls, 2)
            keep = (n.drop_embed if which == 0 else n.drop_core)[t]
            self.step_calls += 1
        assert keep.shape == x.shape, (keep.shape, x.shape)
        return x * keep * (1.0 / (1.0 - p))

    # -- samplers ----------------------------------------------------------------------------
    def sample_gumbel(self, shape, eps=1e-20):
        U = self.noises[max(self.cur, 0)].U[self.t_gumbel]
        self.t_gumbel += 1
        assert tuple(U.shape) == tuple(shape)
        return -torch.log(-torch.log(U + eps) + eps)

    def multinomial(self, p, n, *a, **k):
        E = self.noises[max(self.cur, 0)].E[self.t_mult]
        self.t_mult += 1
        return (p / E).max(dim=1)[1][:, None]

    def __enter__(self):
        import torch.nn.functional as F
        self._saved = (F.dropout, torch.multinomial, torch.Tensor.uniform_,
                       self.m_gumbel.sample_gumbel, self.m_gumbel_soft.sample_gumbel)
        F.dropout = self.dropout
        torch.multinomial = self.multinomial
        self.m_gumbel.sample_gumbel = self.sample_gumbel
        self.m_gumbel_soft.sample_gumbel = self.sample_gumbel
        inj = self
        orig_uniform = torch.Tensor.uniform_

        def uniform_(tensor, *a, **k):
            n = inj.noises[max(inj.cur, 0)]
            u = n.part_u if n.part_u is not None else n.ss_u      # scheduled sampling (AttModel.py:119)
            if u is not None and tensor.dim() == 1 and tensor.numel() == inj.rows:
                tensor.copy_(u[inj.t_part])
                inj.t_part += 1
                return tensor
            return orig_uniform(tensor, *a, **k)

        torch.Tensor.uniform_ = uniform_
        return self

    def __exit__(self, *exc):
        import torch.nn.functional as F
        (F.dropout, torch.multinomial, torch.Tensor.uniform_,
         self.m_gumbel.sample_gumbel, self.m_gumbel_soft.sample_gumbel) = self._saved


def build_reference(ref, dims, Ps, Pl, rows, **opt_over):
    opt = ref_loader.reference_opt(**asdict(dims), batch_size=rows, **opt_over)
    model = ref.AlternatingJointModel(opt)
    sd = {"caption_generator." + k: v.clone() for k, v in Ps.items()}
    sd.update({"vse." + k: v.clone() for k, v in Pl.items()})
    missing = model.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    model.train()
    return model, opt


def grads_of(model):
    out = {}
    for name, p in model.named_parameters():
        if not name.startswith(("caption_generator.", "vse.")):
            continue   # prev_* deep copies made by changeModelUpdateStatus
        out[name] = torch.zeros_like(p) if p.grad is None else p.grad.detach().clone()
    return out


def oracle_grads(loss, Ps, Pl):
    names = ["caption_generator." + k for k in Ps] + ["vse." + k for k in Pl]
    tensors = list(Ps.values()) + list(Pl.values())
    gs = torch.autograd.grad(loss, tensors, allow_unused=True)
    return {n: (torch.zeros_like(t) if g is None else g) for n, t, g in zip(names, tensors, gs)}


def leaf(P):
    return {k: v.clone().requires_grad_(True) for k, v in P.items()}


def compare(tag, a, b, rtol=2e-4, atol=2e-6):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    if a.dtype in (torch.long, torch.int64, torch.bool):
        assert torch.equal(a, b), f"{tag}: integer mismatch"
        return
    scale = max(b.abs().max().item(), 1e-30)
    err = (a - b).abs().max().item()
    assert err <= atol + rtol * scale, f"{tag}: max abs err {err} (scale {scale})"


def run_case(ref, name, dims, *, rows, regions, varlen, mode, kind, tau=1.0, dropout=True,
             baseline="gt", weight=0.01, seed=0, eos_bias=0.0, prob=0.25, ss_prob=0.0,
             sample_max=0, decoding_constraint=0, vse=None, cider=0.0, spi=1, use_gen=0, beam_size=1):
    Ps = synth.speaker_params(dims, seed=seed, eos_bias=eos_bias)
    Pl = synth.listener_params(dims, seed=seed + 1)
    batch = synth.make_batch(dims, rows, regions, seed=seed + 2, varlen=varlen, min_regions=2)
    need_g = mode in ("gumbel", "gumbel_softmax")
    need_m = mode in ("multinomial", "multinomial_soft", "reinforce") or ss_prob > 0 or kind == "decode"
    noise = synth.make_noise(dims, rows, regions, seed + 3, dropout=dropout, gumbel=need_g,
                             multinomial=need_m, partial=mode.endswith("soft") or mode == "gumbel_softmax",
                             sched=ss_prob > 0)
    noise2 = synth.make_noise(dims, rows, regions, seed + 4, dropout=dropout)
    drop_p = 0.5 if dropout else 0.0
    cfg = OJ.JointCfg(vocab_size=dims.vocab_size, seq_length=dims.seq_length, drop_p=drop_p,
                      retrieval_reward=mode, gumbel_temp=tau, multinomial_temp=tau,
                      prob_gumbel_softmax=prob, prob_multinomial_soft=prob,
                      retrieval_reward_weight=weight, reinforce_baseline_type=baseline,
                      vse_loss_weight=1.0 if kind == "listener_turn" else 0.0,
                      caption_loss_weight=1.0 if kind == "mle" else 0.0)
    model, opt = build_reference(
        ref, dims, Ps, Pl, rows, retrieval_reward=mode, gumbel_temp=tau, multinomial_temp=tau,
        drop_prob_lm=drop_p, retrieval_reward_weight=0.0 if kind == "mle" else weight,
        reinforce_baseline_type=baseline, prob_gumbel_softmax=prob, prob_multinomial_soft=prob,
        vse_loss_weight=cfg.vse_loss_weight, caption_loss_weight=cfg.caption_loss_weight,
        is_alternating=0 if kind == "mle" else 1, continue_from_existing_models=False,
        alternating_turn=None if kind == "mle" else ["speaker", "listener"],
        cider_optimization=cider, use_gen_cider_scores=use_gen)
    data = {}
    if cider > 0:
        import misc.rewards as R                      # the reference's own scorer, "corpus" mode
        R.CiderD_scorer = None
        R.init_scorer("corpus")
        data = {"gts": OJ.gts_from_labels(batch.labels, spi, seed)}
    captured = {}
    orig_sample = model.caption_generator.sample

    def rec_sample(*a, **k):
        out = orig_sample(*a, **k)
        captured.setdefault("samples", []).append(out)
        return out

    model.caption_generator.sample = rec_sample
    model.caption_generator.ss_prob = ss_prob
    out = {}
    meta = dict(name=name, dims=asdict(dims), rows=rows, regions=regions, varlen=varlen, mode=mode,
                kind=kind, tau=tau, dropout=dropout, baseline=baseline, weight=weight, seed=seed,
                eos_bias=eos_bias, prob=prob, ss_prob=ss_prob, sample_max=sample_max,
                decoding_constraint=decoding_constraint, vse=vse)
    if cider > 0:
        meta.update(cider=cider, spi=spi, use_gen=use_gen)
    if kind == "beam":
        # AttModel.sample_beam under model.eval() (eval_utils.py:187 style call)
        from oracle import cases as OC
        meta.update(beam_size=beam_size)
        spk = model.caption_generator
        spk.eval()
        with torch.no_grad():
            seq, lp = orig_sample(batch.fc_feats, batch.att_feats, batch.att_masks,
                                  {"beam_size": beam_size, "decoding_constraint": decoding_constraint})
        o = OC.run_oracle(meta)
        compare(name + ".seq", o["seq"], seq)
        compare(name + ".logprobs", o["logprobs"], lp)
        done_n = np.array([len(d) for d in spk.done_beams], np.int64)
        assert done_n.tolist() == [len(d) for d in o["done_beams"]]
        done_p = np.full((rows, int(done_n.max())), np.nan, np.float32)
        done_seq = np.zeros((rows, int(done_n.max()), dims.seq_length), np.int64)
        for k in range(rows):
            for e, (a, b) in enumerate(zip(spk.done_beams[k], o["done_beams"][k])):
                assert torch.equal(a["seq"], b["seq"]) and abs(float(a["p"]) - b["p"]) <= 1e-4
                done_p[k, e], done_seq[k, e] = float(a["p"]), a["seq"].numpy()
        blob = {"meta": np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8),
                "out.seq": seq.numpy(), "out.logprobs": lp.numpy(), "out.done_n": done_n,
                "out.done_p": done_p, "out.done_seq": done_seq}
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **blob)
        print(f"[golden] {name:34s} beam {beam_size} done {done_n.tolist()} min gap {float(o['gaps'].min()):.2e} OK")
        return
    if kind == "vse":
        # VSEFCModel.forward alone, non-default listener options (vse_pool_type / vse_use_abs /
        # vse_max_violation), loss and parameter gradients
        from oracle import cases as OC
        vopt = ref_loader.reference_opt(**asdict(dims), batch_size=rows, vse_pool_type=vse["pool_type"],
                                        vse_use_abs=vse["use_abs"], vse_max_violation=vse["max_violation"])
        vmodel = ref.VSEFCModel(vopt)
        vmodel.load_state_dict({k: v.clone() for k, v in Pl.items()}, strict=True)
        vmodel.train()
        lr = vmodel(batch.fc_feats, None, batch.labels, batch.masks, vse["whole_batch"], vse["only"])
        total = (lr * torch.linspace(0.5, 1.5, lr.numel())).sum() if vse["whole_batch"] else lr
        total.backward()
        o = OC.run_oracle(meta)
        compare(name + ".loss_rows", o["loss_rows"], lr.detach())
        compare(name + ".loss", o["loss"], total.detach())
        blob = {"meta": np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8),
                "out.loss_rows": lr.detach().numpy(), "out.loss": total.detach().numpy()}
        for k, p in vmodel.named_parameters():
            g = torch.zeros_like(p) if p.grad is None else p.grad
            compare(f"{name}.grad[{k}]", o["grads"]["vse." + k], g, rtol=5e-4, atol=1e-7)
            blob["grad.vse." + k] = g.numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **blob)
        print(f"[golden] {name:34s} vse loss={float(total):+.6f} OK")
        return
    if kind == "decode":
        # AttModel.sample with index outputs (eval_utils.py:187 style call), optional constraint
        from oracle import cases as OC
        with Inject(ref, [noise, noise2], rows, batch.att_masks), torch.no_grad():
            seq, lp = orig_sample(batch.fc_feats, batch.att_feats, batch.att_masks,
                                  {"sample_max": sample_max, "temperature": tau,
                                   "decoding_constraint": decoding_constraint})
        o = OC.run_oracle(meta)
        compare(name + ".seq", o["seq"], seq)
        compare(name + ".logprobs", o["logprobs"], lp)
        blob = {"meta": np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8),
                "out.seq": seq.numpy(), "out.logprobs": lp.numpy()}
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **blob)
        rep = int((seq[:, 1:] == seq[:, :-1])[seq[:, 1:] > 0].sum())
        print(f"[golden] {name:34s} decode tokens={int(seq.numel())} repeats={rep} OK")
        return
    with Inject(ref, [noise, noise2], rows, batch.att_masks):
        if kind == "mle":
            loss = model(batch.fc_feats, batch.labels, batch.masks, {}, batch.att_feats,
                         batch.att_masks)
        elif kind == "listener_turn":
            loss = model(batch.fc_feats, batch.labels, batch.masks, {}, batch.att_feats,
                         batch.att_masks, is_alternating=True, alternating_turn="listener")
        else:
            loss = model(batch.fc_feats, batch.labels, batch.masks, data, batch.att_feats,
                         batch.att_masks, is_alternating=True, alternating_turn="speaker")
        loss = loss.sum()
        loss.backward()
    ref_grads = grads_of(model)
    out["loss"] = loss.detach()
    if captured.get("samples"):
        s0 = captured["samples"][0]
        out["seq"] = s0[0].detach()
        out["logprobs"] = s0[-1].detach()
        if len(captured["samples"]) > 1:
            out["seq_greedy"] = captured["samples"][1][0].detach()
    if cider > 0:
        for k in ("loss_cider", "avg_reward", "cider_greedy"):
            out[k] = torch.as_tensor(float(model._loss[k]), dtype=torch.float64)

    # ---- oracle on the same tensors ---------------------------------------------------------
    if cider > 0:
        from oracle import cases as OC
        o = OC.run_oracle(meta)
        for k in ("seq", "logprobs", "seq_greedy", "loss_cider", "avg_reward", "cider_greedy", "loss"):
            compare(f"{name}.{k}", o[k], out[k], rtol=2e-4, atol=2e-6)
        for k in ref_grads:
            compare(f"{name}.grad[{k}]", o["grads"][k], ref_grads[k], rtol=5e-4, atol=1e-7)
        blob = {"meta": np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)}
        for k, v in out.items():
            blob["out." + k] = v.numpy()
        for k, v in ref_grads.items():
            blob["grad." + k] = v.numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **blob)
        print(f"[golden] {name:34s} loss={out['loss'].item():+.6f} loss_cider={float(out['loss_cider']):+.6f} "
              f"avg_reward={float(out['avg_reward']):+.4f} OK")
        return
    Pso, Plo = leaf(Ps), leaf(Pl)
    if kind == "mle":
        fed = []
        o_loss = OJ.mle_loss(Pso, batch.att_feats, batch.att_masks, batch.labels, batch.masks,
                             noise, cfg, ss_prob=ss_prob, fed_out=fed)
        out["fed"] = torch.stack(fed, 1)       # oracle's own record (the reference does not expose it)
        if ss_prob > 0:
            swapped = int((out["fed"][:, 1:] != batch.labels[:, 1: out["fed"].size(1)]).sum())
            assert swapped > 0, "scheduled sampling never replaced an input: pick another seed"
    elif kind == "listener_turn":
        o_loss, res, _ = OJ.listener_turn_loss(Pso, Plo, batch.fc_feats, batch.att_feats,
                                               batch.att_masks, noise, cfg)
        compare(name + ".seq", res.seq, out["seq"])
    elif mode == "reinforce":
        o_loss, res, r, b = OJ.reinforce_speaker_loss(
            Pso, Plo, batch.fc_feats, batch.att_feats, batch.att_masks, batch.labels, batch.masks,
            noise, cfg, noise_greedy=noise2)
        compare(name + ".seq", res.seq, out["seq"])
        compare(name + ".logprobs", res.logprobs.detach(), out["logprobs"])
        out["reward"], out["baseline"] = r.detach(), b.detach()
    else:
        o_loss, res, masks, loss_vse = OJ.st_joint_loss(Pso, Plo, batch.fc_feats, batch.att_feats,
                                                        batch.att_masks, noise, cfg)
        compare(name + ".seq", res.seq, out["seq"])
        compare(name + ".logprobs", res.logprobs.detach(), out["logprobs"])
        out["loss_vse"] = loss_vse.detach()
    compare(name + ".loss", o_loss.detach(), out["loss"])
    o_grads = oracle_grads(o_loss, Pso, Plo)
    for k in ref_grads:
        compare(f"{name}.grad[{k}]", o_grads[k], ref_grads[k], rtol=5e-4, atol=1e-7)

    blob = {"meta": np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)}
    for k, v in out.items():
        blob["out." + k] = v.numpy()
    small = sum(v.numel() for v in Ps.values()) < 200_000
    for k, v in ref_grads.items():
        if small:
            blob["grad." + k] = v.numpy()
        else:  # real dims: keep norms and a strided sample instead of 100 MB of gradients
            blob["gradnorm." + k] = np.array(v.norm().item(), dtype=np.float64)
            blob["gradsample." + k] = v.flatten()[:: max(1, v.numel() // 257)][:257].numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **blob)
    n_tok = int(out["seq"].numel()) if "seq" in out else 0
    print(f"[golden] {name:34s} loss={out['loss'].item():+.6f} tokens={n_tok} OK")


CASES = [
    # name, dims, kwargs
    ("tiny_gumbel", synth.TINY, dict(rows=6, regions=5, varlen=False, mode="gumbel", kind="speaker_turn")),
    ("tiny_gumbel_tau075_masks", synth.TINY, dict(rows=7, regions=6, varlen=True, mode="gumbel",
                                                  kind="speaker_turn", tau=0.75, seed=10)),
    ("tiny_gumbel_nodrop_tau8", synth.TINY, dict(rows=5, regions=4, varlen=False, mode="gumbel",
                                                 kind="speaker_turn", tau=8.0, dropout=False, seed=20)),
    ("tiny_multinomial", synth.TINY, dict(rows=6, regions=5, varlen=True, mode="multinomial",
                                          kind="speaker_turn", seed=30)),
    ("tiny_gumbel_softmax", synth.TINY, dict(rows=6, regions=5, varlen=False, mode="gumbel_softmax",
                                             kind="speaker_turn", seed=40, prob=0.5)),
    ("tiny_multinomial_soft", synth.TINY, dict(rows=6, regions=5, varlen=False, mode="multinomial_soft",
                                               kind="speaker_turn", seed=50, prob=0.5)),
    ("tiny_reinforce_gt", synth.TINY, dict(rows=6, regions=5, varlen=True, mode="reinforce",
                                           kind="speaker_turn", baseline="gt", weight=0.8, seed=60)),
    ("tiny_reinforce_greedy", synth.TINY, dict(rows=6, regions=5, varlen=False, mode="reinforce",
                                               kind="speaker_turn", baseline="greedy", weight=0.8, seed=70)),
    ("tiny_reinforce_none", synth.TINY, dict(rows=6, regions=5, varlen=False, mode="reinforce",
                                             kind="speaker_turn", baseline="no", weight=0.8, seed=80)),
    ("tiny_listener_turn", synth.TINY, dict(rows=6, regions=5, varlen=False, mode="reinforce",
                                            kind="listener_turn", seed=90)),
    ("tiny_mle", synth.TINY, dict(rows=6, regions=5, varlen=True, mode="gumbel", kind="mle", seed=100)),
    ("tiny_mle_sched_sampling", synth.TINY, dict(rows=6, regions=5, varlen=True, mode="gumbel", kind="mle",
                                                 seed=110, ss_prob=0.5)),
    ("tiny_decode_constraint_greedy", synth.TINY, dict(rows=6, regions=5, varlen=True, mode="reinforce",
                                                       kind="decode", seed=120, sample_max=1,
                                                       decoding_constraint=1, dropout=False)),
    ("tiny_decode_constraint_sampled", synth.TINY, dict(rows=6, regions=5, varlen=False, mode="reinforce",
                                                        kind="decode", seed=130, sample_max=0, tau=0.5,
                                                        decoding_constraint=1)),
    ("tiny_vse_mean_abs", synth.TINY, dict(rows=7, regions=3, varlen=False, mode="gumbel", kind="vse", seed=140,
                                           vse=dict(pool_type="mean", use_abs=1, max_violation=1,
                                                    whole_batch=False, only="off"))),
    ("tiny_vse_max_sumviolation", synth.TINY, dict(rows=7, regions=3, varlen=False, mode="gumbel", kind="vse",
                                                   seed=150, vse=dict(pool_type="max", use_abs=0, max_violation=0,
                                                                      whole_batch=True, only="off"))),
    ("tiny_vse_last_sumviolation_image", synth.TINY, dict(rows=6, regions=3, varlen=False, mode="gumbel",
                                                          kind="vse", seed=160,
                                                          vse=dict(pool_type="last", use_abs=0, max_violation=0,
                                                                   whole_batch=False, only="image"))),
    ("tiny_gumbel_cider", synth.TINY, dict(rows=6, regions=5, varlen=True, mode="gumbel", kind="speaker_turn",
                                           seed=300, cider=0.5, spi=2)),
    ("tiny_reinforce_gt_cider", synth.TINY, dict(rows=6, regions=5, varlen=False, mode="reinforce",
                                                 kind="speaker_turn", baseline="gt", weight=0.8, seed=310,
                                                 cider=0.2, spi=3)),
    ("tiny_reinforce_greedy_cider_gen", synth.TINY, dict(rows=6, regions=5, varlen=False, mode="reinforce",
                                                         kind="speaker_turn", baseline="greedy", weight=0.8,
                                                         seed=320, cider=1.0, spi=2, use_gen=1)),
    ("tiny_cider_only", synth.TINY, dict(rows=8, regions=4, varlen=True, mode="reinforce", kind="speaker_turn",
                                         weight=0.0, seed=330, cider=1.0, spi=2)),
    ("tiny_beam3", synth.TINY, dict(rows=5, regions=4, varlen=True, mode="reinforce", kind="beam", seed=400,
                                    beam_size=3, dropout=False)),
    ("tiny_beam2_constraint", synth.TINY, dict(rows=4, regions=5, varlen=False, mode="reinforce", kind="beam",
                                               seed=410, beam_size=2, decoding_constraint=1, eos_bias=1.0,
                                               dropout=False)),
    ("real_beam2_eos", synth.Dims(), dict(rows=3, regions=6, varlen=True, mode="reinforce", kind="beam",
                                          seed=420, beam_size=2, eos_bias=4.0, dropout=False)),
    ("real_beam3_constraint", synth.Dims(), dict(rows=3, regions=6, varlen=False, mode="reinforce", kind="beam",
                                                 seed=430, beam_size=3, decoding_constraint=1, dropout=False)),
    ("real_gumbel_b4", synth.Dims(), dict(rows=4, regions=6, varlen=True, mode="gumbel",
                                          kind="speaker_turn", seed=200, eos_bias=6.0)),
    ("real_mle_b4", synth.Dims(), dict(rows=4, regions=6, varlen=False, mode="gumbel", kind="mle",
                                       seed=210)),
    ("real_reinforce_gt_b4", synth.Dims(), dict(rows=4, regions=6, varlen=True, mode="reinforce",
                                                kind="speaker_turn", baseline="gt", weight=0.8, seed=220,
                                                eos_bias=6.0)),
    ("real_multinomial_soft_b4", synth.Dims(), dict(rows=4, regions=6, varlen=False, mode="multinomial_soft",
                                                    kind="speaker_turn", seed=230, eos_bias=6.0, prob=0.5,
                                                    tau=0.75)),
    ("real_listener_turn_b4", synth.Dims(), dict(rows=4, regions=6, varlen=True, mode="reinforce",
                                                 kind="listener_turn", seed=240, eos_bias=6.0)),
]


def main():
    ref = ref_loader.load_reference()
    only = set(sys.argv[1:])
    for name, dims, kw in CASES:
        if only and name not in only:
            continue
        run_case(ref, name, dims, **kw)


if __name__ == "__main__":
    main()
