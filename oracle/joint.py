"""TEST INFRASTRUCTURE ONLY (checker, never the product path).

CPU fp32 restatement of the loss recipes of AlternatingJointModel for the hot path: the
straight-through / partial-sampling joint loss (`st_and_ps_methods`), REINFORCE with the listener
reward and its baselines, the listener turn on generated captions and the MLE/VSE terms.
The CIDEr self-critical term (`traditional_cider`) uses the scorer restatement in oracle/cider.py.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

import numpy as np

from . import cider as CD
from . import listener as L
from . import speaker as S


@dataclass
class JointCfg:
    vocab_size: int = 9487
    seq_length: int = 16
    drop_p: float = 0.5
    retrieval_reward: str = "gumbel"
    gumbel_temp: float = 1.0
    multinomial_temp: float = 1.0
    prob_gumbel_softmax: float = 0.25
    prob_multinomial_soft: float = 0.25
    retrieval_reward_weight: float = 0.01
    vse_loss_weight: float = 0.0
    caption_loss_weight: float = 0.0
    reinforce_baseline_type: str = "gt"
    only_one_retrieval: str = "off"
    margin: float = 0.2
    max_violation: bool = True
    pool_type: str = "last"


def caption_masks(word_index):
    """_masks = [1, 1, (w_1 > 0), ..., (w_{n-1} > 0)]     (AlternatingJointModel.py:232-234,353-355)"""
    B = word_index.size(0)
    return torch.cat([torch.ones(B, 2), (word_index > 0).float()[:, :-1]], 1)


def st_joint_loss(Ps, Pl, fc_feats, att_feats, att_masks, noise: S.SpeakerNoise, cfg: JointCfg,
                  forced_tokens: Optional[torch.Tensor] = None, keep_all_steps: bool = False,
                  hinge_replay: Optional[dict] = None):
    """Speaker turn in gumbel / multinomial / *_soft mode: st_and_ps_methods
    (AlternatingJointModel.py:343-376) with VSE weight forced to 0 (:516-518)."""
    V = cfg.vocab_size
    res = S.sample(Ps, att_feats, att_masks, mode=cfg.retrieval_reward, seq_length=cfg.seq_length,
                   vocab_size=V, noise=noise, drop_p=cfg.drop_p, sample_max=0, use_one_hot=1,
                   temperature=1.0, gumbel_temp=cfg.gumbel_temp,
                   multinomial_temp=cfg.multinomial_temp,
                   prob_gumbel_softmax=cfg.prob_gumbel_softmax,
                   prob_multinomial_soft=cfg.prob_multinomial_soft,
                   forced_tokens=forced_tokens, keep_all_steps=keep_all_steps)       # :346-348
    masks = caption_masks(res.seq)                                                   # :353-355
    B = res.seq.size(0)
    bos = torch.zeros(B, 1, V + 2)
    bos[:, 0, V + 1] = 1.0                                                           # :360-369
    seqs = torch.cat([bos, res.one_hots], 1)                                         # :370
    loss_vse = L.vse_forward(Pl, fc_feats, seqs, masks, False, cfg.only_one_retrieval,
                             cfg.margin, cfg.max_violation, cfg.pool_type,
                             hinge_replay)                                           # :371-373
    loss = loss_vse * cfg.retrieval_reward_weight                                    # :374
    return loss, res, masks, loss_vse


def mle_loss(Ps, att_feats, att_masks, seq, masks, noise, cfg: JointCfg, ss_prob: float = 0.0,
             forced_fed=None, fed_out=None):
    """ce_loss -> AttModel.forward                           (AlternatingJointModel.py:196-207)"""
    return S.forward_xe(Ps, att_feats, att_masks, seq, masks, noise=noise, drop_p=cfg.drop_p,
                        ss_prob=ss_prob, forced_fed=forced_fed, fed_out=fed_out)


def vse_gt_loss(Pl, fc_feats, seq, masks, cfg: JointCfg):
    """vse_loss on given (ground-truth or generated index) captions      (:209-224)"""
    return L.vse_forward(Pl, fc_feats, seq, masks, False, cfg.only_one_retrieval, cfg.margin,
                         cfg.max_violation, cfg.pool_type)


def reinforce_speaker_loss(Ps, Pl, fc_feats, att_feats, att_masks, seq_gt, masks_gt,
                           noise: S.SpeakerNoise, cfg: JointCfg,
                           noise_greedy: Optional[S.SpeakerNoise] = None,
                           forced_tokens: Optional[torch.Tensor] = None,
                           forced_tokens_greedy: Optional[torch.Tensor] = None,
                           keep_all_steps: bool = False):
    """Speaker turn, retrieval_reward = 'reinforce' (:226-247,:300-332,:456-481).
    The listener is frozen in this turn (requires_grad False, :571-633)."""
    V = cfg.vocab_size
    res = S.sample(Ps, att_feats, att_masks, mode="reinforce", seq_length=cfg.seq_length,
                   vocab_size=V, noise=noise, drop_p=cfg.drop_p, sample_max=0, temperature=1.0,
                   forced_tokens=forced_tokens, keep_all_steps=keep_all_steps)       # :228-230
    _masks = caption_masks(res.seq)                                                  # :232-234
    B = res.seq.size(0)
    _seqs = torch.cat([torch.full((B, 1), V + 1, dtype=torch.long), res.seq], 1)     # :238-240
    with torch.no_grad():
        r = L.vse_forward(Pl, fc_feats, _seqs, _masks, True, cfg.only_one_retrieval, cfg.margin,
                          cfg.max_violation, cfg.pool_type)                          # :242-244
        if cfg.reinforce_baseline_type == "gt":
            b = L.vse_forward(Pl, fc_feats, seq_gt, masks_gt, True, cfg.only_one_retrieval,
                              cfg.margin, cfg.max_violation, cfg.pool_type)          # :303-304
        elif cfg.reinforce_baseline_type == "greedy":
            g = S.sample(Ps, att_feats, att_masks, mode="reinforce", seq_length=cfg.seq_length,
                         vocab_size=V, noise=noise_greedy or S.SpeakerNoise(), drop_p=cfg.drop_p,
                         sample_max=1, temperature=1.0, forced_tokens=forced_tokens_greedy,
                         keep_all_steps=keep_all_steps)                              # :255-266
            mg = caption_masks(g.seq)
            sg = torch.cat([torch.full((B, 1), V + 1, dtype=torch.long), g.seq], 1)
            b = L.vse_forward(Pl, fc_feats, sg, mg, True, cfg.only_one_retrieval, cfg.margin,
                              cfg.max_violation, cfg.pool_type)                      # :287-290
        else:
            b = torch.zeros_like(r)                                                  # :314
    sc = res.logprobs * (r - b).detach()[:, None] * _masks[:, 1:]                    # :305-309
    sc_loss = sc.sum() / _masks[:, 1:].sum()                                         # :324
    return cfg.retrieval_reward_weight * sc_loss, res, r, b                          # :325


def listener_turn_loss(Ps, Pl, fc_feats, att_feats, att_masks, noise, cfg: JointCfg,
                       forced_tokens=None, keep_all_steps=False):
    """Listener turn (:528-555): sample index captions with a frozen speaker, train the listener
    with the scalar contrastive loss on them, weight vse_loss_weight."""
    V = cfg.vocab_size
    with torch.no_grad():
        res = S.sample(Ps, att_feats, att_masks, mode="reinforce", seq_length=cfg.seq_length,
                       vocab_size=V, noise=noise, drop_p=cfg.drop_p, sample_max=0, temperature=1.0,
                       forced_tokens=forced_tokens, keep_all_steps=keep_all_steps)   # :539-541
    _masks = caption_masks(res.seq)                                                  # :542-544
    B = res.seq.size(0)
    _seqs = torch.cat([torch.full((B, 1), V + 1, dtype=torch.long), res.seq], 1)     # :545-547
    loss_vse = vse_gt_loss(Pl, fc_feats, _seqs, _masks, cfg)
    return cfg.vse_loss_weight * loss_vse, res, loss_vse


def cider_term(sample_logprobs, gen_seq, greedy_seq, gts, *, use_gen_cider_scores: int = 0,
               doc_freq=None, ref_len=None):
    """traditional_cider (AlternatingJointModel.py:407-431): returns (loss_cider, reward [B] float64,
    cider_greedy).  gen_seq / greedy_seq int [B, n]; sample_logprobs [B, n] with autograd;
    gts: list over images of int arrays [n_captions, 16]."""
    gen = gen_seq.detach().cpu().numpy()
    gr = greedy_seq.detach().cpu().numpy()
    cider_gen, diff, cider_greedy = CD.self_critical_reward(gts, gen, gr, doc_freq, ref_len)
    reward = diff if use_gen_cider_scores == 0 else cider_gen                     # :412-419
    gen_masks = caption_masks(gen_seq)                                               # :385-387
    r = torch.from_numpy(-reward.astype("float32"))
    loss_cider = (sample_logprobs * r.unsqueeze(1) * gen_masks[:, 1:]).sum() / gen_masks[:, 1:].sum()
    return loss_cider, reward, cider_greedy


def greedy_for_cider(Ps, att_feats, att_masks, noise, cfg: JointCfg, forced_tokens=None,
                     keep_all_steps=False):
    """greedy_res_for_cider (:391-405): greedy decode in the module's current (training) mode."""
    with torch.no_grad():
        g = S.sample(Ps, att_feats, att_masks, mode="reinforce", seq_length=cfg.seq_length,
                     vocab_size=cfg.vocab_size, noise=noise, drop_p=cfg.drop_p, sample_max=1,
                     temperature=1.0, forced_tokens=forced_tokens, keep_all_steps=keep_all_steps)
    return g


def gts_from_labels(labels, spi: int, extra_seed: int = 0):
    """data['gts'] of a synthetic batch: image i owns the captions of rows i*spi .. (i+1)*spi-1
    (columns 1..T of the label rows, dataloader.py:199-203) plus, for every other image, one more
    seeded random caption (real images have more captions than seq_per_img)."""
    lab = labels.cpu().numpy()
    rows, T = lab.shape[0], lab.shape[1] - 2
    assert rows % spi == 0
    rng = np.random.default_rng(1000 + extra_seed)
    vmax = int(lab.max())
    gts = []
    for i in range(rows // spi):
        caps = [lab[r, 1:T + 1] for r in range(i * spi, (i + 1) * spi)]
        if i % 2 == 0:
            extra = np.zeros(T, lab.dtype)
            k = int(rng.integers(2, T))
            extra[:k] = rng.integers(1, max(vmax, 2) + 1, size=k)
            caps.append(extra)
        gts.append(np.stack(caps, 0))
    return gts
