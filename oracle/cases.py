"""TEST INFRASTRUCTURE ONLY.  Rebuild a golden case from its metadata and run the oracle on it.

Shared by tests/test_oracle_golden.py (oracle vs. stored reference outputs, CPU) and the GPU
parity tests (CUDA path vs. oracle on the same seeded tensors).
"""
from __future__ import annotations

import json
import os
from typing import Dict

import numpy as np
import torch

from . import joint as OJ
from . import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                          "tests", "golden")


def load_golden(name: str):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    return meta, z


def golden_names():
    """Joint / decode / listener cases (make_golden.py); the scorer-only CIDEr-D cases
    (make_golden_cider.py, `cider_*.npz`) are listed by cider_golden_names()."""
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR)
                  if f.endswith(".npz") and not f.startswith(("cider_", "retrieval_")))


def retrieval_golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and f.startswith("retrieval_"))


def cider_golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and f.startswith("cider_"))


def load_cider_golden(name: str):
    """(meta, gts list, gen [B,16], greedy [B,16], doc_freq dict or None, ref_len or None, npz)."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    off = z["gts_off"]
    gts = [z["gts"][off[i]:off[i + 1]] for i in range(len(off) - 1)]
    df = ref_len = None
    if "df_keys" in z.files:
        df = {tuple(int(x) for x in k if x >= 0): float(v) for k, v in zip(z["df_keys"], z["df_val"])}
        ref_len = float(z["ref_len"])
    return meta, gts, z["gen"], z["greedy"], df, ref_len, z


def build_case(meta: Dict):
    dims = synth.Dims(**meta["dims"])
    seed, mode = meta["seed"], meta["mode"]
    rows, regions = meta["rows"], meta["regions"]
    Ps = synth.speaker_params(dims, seed=seed, eos_bias=meta["eos_bias"])
    Pl = synth.listener_params(dims, seed=seed + 1)
    batch = synth.make_batch(dims, rows, regions, seed=seed + 2, varlen=meta["varlen"],
                             min_regions=2)
    noise = synth.make_noise(
        dims, rows, regions, seed + 3, dropout=meta["dropout"],
        gumbel=mode in ("gumbel", "gumbel_softmax"),
        multinomial=mode in ("multinomial", "multinomial_soft", "reinforce") or
        meta.get("ss_prob", 0.0) > 0 or meta["kind"] == "decode",
        partial=mode in ("gumbel_softmax", "multinomial_soft"), sched=meta.get("ss_prob", 0.0) > 0)
    noise2 = synth.make_noise(dims, rows, regions, seed + 4, dropout=meta["dropout"])
    kind = meta["kind"]
    cfg = OJ.JointCfg(
        vocab_size=dims.vocab_size, seq_length=dims.seq_length,
        drop_p=0.5 if meta["dropout"] else 0.0, retrieval_reward=mode, gumbel_temp=meta["tau"],
        multinomial_temp=meta["tau"], prob_gumbel_softmax=meta["prob"],
        prob_multinomial_soft=meta["prob"], retrieval_reward_weight=meta["weight"],
        reinforce_baseline_type=meta["baseline"],
        vse_loss_weight=1.0 if kind == "listener_turn" else 0.0,
        caption_loss_weight=1.0 if kind == "mle" else 0.0)
    return dims, Ps, Pl, batch, noise, noise2, cfg


def run_oracle(meta: Dict, forced_tokens=None, keep_all_steps=False, forced_tokens_greedy=None):
    """Returns dict(loss, seq, logprobs, ..., grads={name: tensor})."""
    dims, Ps, Pl, batch, noise, noise2, cfg = build_case(meta)
    Pso = {k: v.clone().requires_grad_(True) for k, v in Ps.items()}
    Plo = {k: v.clone().requires_grad_(True) for k, v in Pl.items()}
    kind, mode = meta["kind"], meta["mode"]
    out = {}
    if kind == "decode":
        from . import speaker as OS
        with torch.no_grad():
            res = OS.sample(Ps, batch.att_feats, batch.att_masks, mode=mode,
                            seq_length=dims.seq_length, vocab_size=dims.vocab_size, noise=noise,
                            drop_p=cfg.drop_p, sample_max=meta["sample_max"], use_one_hot=0,
                            temperature=meta["tau"],
                            decoding_constraint=meta.get("decoding_constraint", 0))
        return dict(seq=res.seq, logprobs=res.logprobs, grads={})
    if kind == "beam":
        from . import speaker as OS
        with torch.no_grad():
            res = OS.sample_beam(Ps, batch.att_feats, batch.att_masks, seq_length=dims.seq_length,
                                 vocab_size=dims.vocab_size, beam_size=meta["beam_size"],
                                 decoding_constraint=meta.get("decoding_constraint", 0))
        done_n = torch.tensor([len(d) for d in res.done_beams])
        w = int(done_n.max())
        done_p = torch.full((len(res.done_beams), w), float("nan"))
        done_seq = torch.zeros(len(res.done_beams), w, dims.seq_length, dtype=torch.long)
        for k, dl in enumerate(res.done_beams):
            for e, ent in enumerate(dl):
                done_p[k, e], done_seq[k, e] = ent["p"], ent["seq"]
        return dict(seq=res.seq, logprobs=res.logprobs, done_beams=res.done_beams, done_n=done_n,
                    done_p=done_p, done_seq=done_seq, parents=res.parents, toks=res.toks,
                    gaps=res.gaps, grads={})
    if kind == "vse":
        from . import listener as OL
        v = meta["vse"]
        loss = OL.vse_forward(Plo, batch.fc_feats, batch.labels, batch.masks, v["whole_batch"],
                              v["only"], 0.2, bool(v["max_violation"]), v["pool_type"],
                              use_abs=bool(v["use_abs"]))
        out["loss_rows"] = loss.detach()
        if v["whole_batch"]:
            loss = (loss * torch.linspace(0.5, 1.5, loss.numel())).sum()
    elif kind == "mle":
        fed = []
        loss = OJ.mle_loss(Pso, batch.att_feats, batch.att_masks, batch.labels, batch.masks,
                           noise, cfg, ss_prob=meta.get("ss_prob", 0.0), fed_out=fed)
        out["fed"] = torch.stack(fed, 1)
    elif kind == "listener_turn":
        loss, res, _ = OJ.listener_turn_loss(Pso, Plo, batch.fc_feats, batch.att_feats,
                                             batch.att_masks, noise, cfg, forced_tokens,
                                             keep_all_steps)
        out.update(seq=res.seq, logprobs=res.logprobs.detach())
    elif mode == "reinforce" and meta["weight"] == 0 and meta.get("cider", 0) > 0:
        # CIDEr term alone: gen_result_for_cider samples index captions (:378-389,492-496)
        from . import speaker as OS
        res = OS.sample(Pso, batch.att_feats, batch.att_masks, mode="reinforce",
                        seq_length=dims.seq_length, vocab_size=dims.vocab_size, noise=noise,
                        drop_p=cfg.drop_p, sample_max=0, temperature=1.0, forced_tokens=forced_tokens,
                        keep_all_steps=keep_all_steps)
        loss = 0.0
        out.update(seq=res.seq, logprobs=res.logprobs.detach())
    elif mode == "reinforce":
        loss, res, r, b = OJ.reinforce_speaker_loss(
            Pso, Plo, batch.fc_feats, batch.att_feats, batch.att_masks, batch.labels, batch.masks,
            noise, cfg, noise_greedy=noise2, forced_tokens=forced_tokens,
            keep_all_steps=keep_all_steps)
        out.update(seq=res.seq, logprobs=res.logprobs.detach(), reward=r, baseline=b)
    else:
        loss, res, masks, loss_vse = OJ.st_joint_loss(Pso, Plo, batch.fc_feats, batch.att_feats,
                                                      batch.att_masks, noise, cfg, forced_tokens,
                                                      keep_all_steps)
        out.update(seq=res.seq, logprobs=res.logprobs.detach(), loss_vse=loss_vse.detach(),
                   sample=res)
    if meta.get("cider", 0) > 0:
        gts = OJ.gts_from_labels(batch.labels, meta["spi"], meta["seed"])
        g = OJ.greedy_for_cider(Ps, batch.att_feats, batch.att_masks, noise2, cfg,
                                forced_tokens_greedy, keep_all_steps)
        # with keep_all_steps the oracle records all T steps; the reference's caption width is the
        # largest number of leading non-zero ids (AttModel.py:404-408)
        n = max(1, int((res.seq > 0).sum(1).max())) if keep_all_steps else res.seq.size(1)
        loss_cider, reward, cider_greedy = OJ.cider_term(
            res.logprobs[:, :n], res.seq[:, :n], g.seq, gts,
            use_gen_cider_scores=meta.get("use_gen", 0))
        loss = loss + meta["cider"] * loss_cider
        out.update(loss_cider=loss_cider.detach(), avg_reward=torch.tensor(reward.mean()),
                   cider_greedy=torch.tensor(cider_greedy), seq_greedy=g.seq,
                   cider_reward=torch.from_numpy(reward), gts=gts)
    out["loss"] = loss.detach()
    names = ["caption_generator." + k for k in Pso] + ["vse." + k for k in Plo]
    tensors = list(Pso.values()) + list(Plo.values())
    gs = torch.autograd.grad(loss, tensors, allow_unused=True)
    out["grads"] = {n: (torch.zeros_like(t) if g is None else g)
                    for n, t, g in zip(names, tensors, gs)}
    return out
