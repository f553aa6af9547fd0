"""Drop-in joint speaker-listener model (reference: models/AlternatingJointModel.py).

Keeps the reference's constructor, `forward(fc_feats, seq, masks, data, att_feats, att_masks,
is_alternating, alternating_turn)` turn logic (:433-555), the loss recipes of the hot path
(`ce_loss`, `vse_loss`, `reinforce_disc`, `gt/greedy/no_baseline`, `loss_configuration`,
`st_and_ps_methods`, `gen_result_for_cider`, `greedy_res_for_cider`, `traditional_cider`), `sample`,
`loss()`, `get/setLossFlages`.  The CIDEr-D self-critical reward (misc/rewards.py) is scored on the
device (rewards.py / csrc/cider.cu): captions never travel to the host.

The straight-through joint step (`retrieval_reward` in {'gumbel','multinomial'}) runs as ONE fused
autograd node: speaker decode -> listener loss forward; listener backward -> factored
straight-through gradient (demb . W_emb^T formed tile by tile, never a dense [B,n,V+2] tensor) ->
speaker BPTT.  The reference hands a dense one-hot tensor across this boundary
(AttModel.py:445-452 -> VSEFCModel.py:104); that path still exists (speaker.sample with
use_one_hot=1 + vse(one_hots)) for callers that want the tensors.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from .. import engine as EN
from .. import rewards
from .AttModel import _SpeakerLossFn
from .AttModel import _ordered as _ordered_s


def _setup(*a, **k):
    from . import setup
    return setup(*a, **k)


class _StJointFn(torch.autograd.Function):
    """loss_vse of st_and_ps_methods (:343-376) as a function of speaker + listener parameters.
    Second output: the sampled ids' log-probabilities [n_steps, B] (sampleLogprobs, AttModel.py:
    346,357), which the CIDEr term differentiates through the same pass (:490-503); when nothing
    consumes them their gradient arrives as None and costs nothing."""

    @staticmethod
    def forward(ctx, owner, sp, lp, *params):
        ctx.owner, ctx.sp, ctx.lp = owner, EN.retain(sp), EN.retain(lp)
        ctx.set_materialize_grads(False)
        return lp.t["loss"][0].clone(), sp.t["logp"][: sp.n_steps].clone()

    @staticmethod
    def backward(ctx, g, g_logp=None):
        owner, sp, lp = ctx.owner, ctx.sp, ctx.lp
        if g is None:
            g = torch.zeros(1, device=lp.t["loss"].device)
        spk, lis = owner.caption_generator, owner.vse
        Ps, Pl = spk._params(), lis._params()
        need_l = any(p.requires_grad for p in Pl.values())
        need_s = any(p.requires_grad for p in Ps.values())
        # gradients of bucketed parameters are written straight into the optimizer's flat bucket
        Gl_t, dl = EN.grad_targets(Pl, EN.LISTENER_PARAM_NAMES) if need_l else (None, set())
        Gl, demb16 = EN.listener_backward(lp, Pl, g_loss=g.contiguous().float().reshape(1),
                                          need_param_grads=need_l, out=Gl_t)
        if need_l:
            EN.adopt_direct(Pl, Gl, dl)
            # the listener's gradients are final here (unless a parameter is shared with the
            # speaker): exchange them between the ranks now, hidden behind the speaker's BPTT
            mine = {id(p) for p in Ps.values()}
            final = [Pl[n] for n in dl if id(Pl[n]) not in mine]
            buckets = {id(getattr(p, "_coopcap_bucket", None)): p._coopcap_bucket for p in final}
            if len(final) == len(dl) == len(EN.LISTENER_PARAM_NAMES) and len(buckets) == 1:
                next(iter(buckets.values())).reduce_async(final)
        gs = (None,) * len(EN.SPEAKER_PARAM_NAMES)
        if need_s:
            T = sp.n_steps
            dz16 = EN.st_logit_grads(sp, demb16[1:T + 1], lis._packed.get(Pl)["w_emb16"])
            if g_logp is not None:
                EN.logp_logit_grads(sp, sp.t["tok_fed"][1:T + 1].contiguous(),
                                    g_logp.contiguous().float(), into=dz16)
            Gs_t, ds = EN.grad_targets(Ps, EN.SPEAKER_PARAM_NAMES)

            def logit_grads_ready(G):
                # 19.5 MB of the speaker's 57.8 MB are final after the vocabulary layer's backward
                names = ("logit.weight", "logit.bias")
                if all(n in ds for n in names):
                    b = Ps[names[0]]._coopcap_bucket
                    if all(getattr(Ps[n], "_coopcap_bucket", None) is b for n in names):
                        b.reduce_async([Ps[n] for n in names])

            Gs = EN.speaker_backward(sp, dz16, Ps, out=Gs_t, after_logit_layer=logit_grads_ready)
            EN.adopt_direct(Ps, Gs, ds)
            gs = tuple((None if n in ds else Gs[n].view_as(Ps[n])) for n in EN.SPEAKER_PARAM_NAMES)
        gl = tuple((None if (not need_l or n in dl) else Gl[n].view_as(Pl[n]))
                   for n in EN.LISTENER_PARAM_NAMES)
        EN.release(sp)
        EN.release(lp)
        ctx.sp = ctx.lp = None
        return (None, None, None) + gs + gl


class _PsJointFn(torch.autograd.Function):
    """loss_vse of st_and_ps_methods for the partial-sampling modes (gumbel_softmax /
    multinomial_soft): the listener embeds the emitted vectors with a dense contraction and the
    speaker's BPTT forms d(loss)/d(logits) step by step (the vectors also feed the next input)."""

    @staticmethod
    def forward(ctx, owner, sp, lp, *params):
        ctx.owner, ctx.sp, ctx.lp = owner, EN.retain(sp), EN.retain(lp)
        return lp.t["loss"][0].clone()

    @staticmethod
    def backward(ctx, g):
        owner, sp, lp = ctx.owner, ctx.sp, ctx.lp
        spk, lis = owner.caption_generator, owner.vse
        Ps, Pl = spk._params(), lis._params()
        need_l = any(p.requires_grad for p in Pl.values())
        need_s = any(p.requires_grad for p in Ps.values())
        Gl, demb16 = EN.listener_backward(lp, Pl, g_loss=g.contiguous().float().reshape(1),
                                          need_param_grads=need_l)
        T = sp.n_steps
        if need_l:
            EN.caption_embed_dense_bwd(sp.t["soft16"], demb16, T, spk.vocab_size + 1,
                                       Gl["txt_enc.embed.weight"])
        gs = (None,) * len(EN.SPEAKER_PARAM_NAMES)
        if need_s:
            Gs = EN.speaker_backward(sp, None, Ps, ps_demb16=demb16[1:T + 1],
                                     ps_w_emb16=lis._packed.get(Pl)["w_emb16"])
            gs = tuple(Gs[n].view_as(Ps[n]) for n in EN.SPEAKER_PARAM_NAMES)
        gl = tuple((Gl[n].view_as(Pl[n]) if need_l else None) for n in EN.LISTENER_PARAM_NAMES)
        EN.release(sp)
        EN.release(lp)
        ctx.sp = ctx.lp = None
        return (None, None, None) + gs + gl


class AlternatingJointModel(nn.Module):
    def __init__(self, opt, iteration=None):
        super().__init__()
        self.opt = opt
        self.use_word_weights = getattr(opt, "use_word_weights", 0)
        self.caption_generator = _setup(opt, opt.caption_model, "caption_model")       # :78
        if opt.vse_model != "None":
            self.vse = _setup(opt, opt.vse_model, "vse_model")                         # :82
            self.share_embed = opt.share_embed
            if self.share_embed:
                self.caption_generator.embed[0] = self.vse.txt_enc.embed               # :85-88
                if self.opt.phase == 2:
                    for p in self.caption_generator.embed.parameters():
                        p.requires_grad = False
        else:
            raise NotImplementedError("vse_model='None' (speaker only) : use models.setup directly")
        if opt.retrieval_reward == "reinforce":
            if opt.vse_loss_weight == 0:
                for p in self.vse.parameters():
                    p.requires_grad = False                                            # :96-99
        self.batch_size = opt.batch_size
        self.vse_loss_weight = opt.vse_loss_weight
        self.caption_loss_weight = opt.caption_loss_weight
        self.retrieval_reward = opt.retrieval_reward
        if getattr(opt, "alternating_turn", None) is not None:
            if len(opt.alternating_turn) == 1 and opt.retrieval_reward == "reinforce":
                if opt.alternating_turn[0] == "listener":
                    opt.retrieval_reward_weight = 0                                    # :110-114
        self.retrieval_reward_weight = opt.retrieval_reward_weight
        self.reinforce_baseline_type = getattr(opt, "reinforce_baseline_type", "greedy")
        self.only_one_retrieval = getattr(opt, "only_one_retrieval", "off")
        self.cider_optimization = getattr(opt, "cider_optimization", 0)
        self.use_gen_cider_scores = getattr(opt, "use_gen_cider_scores", 0)
        self._loss = {}
        self._load_checkpoints(opt, iteration)

    # checkpoint loading (:131-177): same file names and precedence, tolerant state-dict loader
    def _load_checkpoints(self, opt, iteration):
        from . import load, load_state_dict
        if getattr(opt, "is_alternating", 0):
            if getattr(opt, "continue_from_existing_models", False):
                # a previous alternating snapshot wins; else the speaker pre-trained in stage 2
                start = vars(opt).get("start_from", None)
                path = os.path.join(start, "alternatingModel.pth") if start is not None else None
                if path is not None and os.path.isfile(path):
                    if iteration:
                        path = os.path.join(start, "alternatingModel-" + str(iteration) + ".pth")
                    print("Loaded alternating model from {}".format(path))
                else:
                    path = getattr(opt, "speaker_stage_2_model_path", None)
                    if path is None:        # nothing to resume from (the reference would fail here)
                        return
                    print(f'Loaded pre-trained "speaker" model, after stage 2 from {path}')
                load_state_dict(self, torch.load(path, map_location="cpu"))
            return
        load(self, opt, iteration)
        init = getattr(opt, "initialize_retrieval", None)
        if init is not None:
            sd = torch.load(init, map_location="cpu")
            load_state_dict(self, {k: v for k, v in sd.items() if "vse." in k})

    # ------------------------------------------------------------------ flags (:180-194)
    def getLossFlags(self):
        return [self.vse_loss_weight, self.caption_loss_weight, self.cider_optimization,
                self.retrieval_reward_weight]

    def setLossFlages(self, VSEWeight, MLEWeight, ciderFlag, DISCWeight):
        self.vse_loss_weight = VSEWeight
        self.caption_loss_weight = MLEWeight
        self.cider_optimization = ciderFlag
        self.retrieval_reward_weight = DISCWeight

    def _zero(self, ref):
        return torch.zeros(1, device=ref.device)

    # ------------------------------------------------------------------ loss terms
    def ce_loss(self, fc_feats, att_feats, att_masks, seq, masks):                     # :196-207
        if self.caption_loss_weight > 0:
            loss_cap = self.caption_generator(fc_feats, att_feats, att_masks, seq, masks)
            self._loss["loss_cap"] = loss_cap.detach()
            return loss_cap
        return self._zero(fc_feats)

    def vse_loss(self, fc_feats, att_feats, seq, masks, only_one_retrieval):            # :209-224
        if self.vse_loss_weight > 0:
            loss_vse = self.vse(fc_feats, att_feats, seq, masks,
                                only_one_retrieval=self.only_one_retrieval)
            self._loss["loss_vse"] = loss_vse.detach()
            return loss_vse
        return self._zero(fc_feats)

    @staticmethod
    def _caption_masks(seqs):
        """[1, 1, (w_1>0), ..., (w_{n-1}>0)]                                  (:232-234,353-355)"""
        B = seqs.size(0)
        return torch.cat([torch.ones(B, 2, device=seqs.device), (seqs > 0).float()[:, :-1]], 1)

    def _with_bos(self, seqs):
        bos = torch.full((seqs.size(0), 1), self.caption_generator.vocab_size + 1,
                         dtype=seqs.dtype, device=seqs.device)
        return torch.cat([bos, seqs], 1)

    def reinforce_disc(self, fc_feats, att_feats, att_masks):                           # :226-247
        _seqs, _sampleLogProbs = self.caption_generator.sample(
            fc_feats, att_feats, att_masks, {"sample_max": 0, "temperature": 1})
        gen_result, sample_logprobs = _seqs, _sampleLogProbs
        _masks = self._caption_masks(_seqs)
        gen_masks = _masks
        _seqs = self._with_bos(_seqs)
        retrieval_loss = self.vse(fc_feats, att_feats, _seqs, _masks, True,
                                  only_one_retrieval=self.only_one_retrieval)
        return retrieval_loss, _seqs, _masks, gen_result, sample_logprobs, gen_masks

    def greedy_baseline(self, fc_feats, att_feats, att_masks, retrieval_loss, _seqs,
                        _sampleLogProbs, _masks):                                       # :250-298
        with torch.no_grad():
            _seqs_greedy, _ = self.caption_generator.sample(
                fc_feats, att_feats, att_masks, opt={"sample_max": 1, "temperature": 1})
            greedy_res = _seqs_greedy
            _masks_greedy = self._caption_masks(_seqs_greedy)
            baseline = self.vse(fc_feats, att_feats, self._with_bos(_seqs_greedy), _masks_greedy,
                                True, only_one_retrieval=self.only_one_retrieval)
        sc_loss = _sampleLogProbs * (retrieval_loss - baseline).detach().unsqueeze(1) * \
            _masks[:, 1:].detach().float()
        return baseline, sc_loss, greedy_res

    def gt_baseline(self, fc_feats, att_feats, att_masks, retrieval_loss, _seqs, _sampleLogProbs,
                    _masks, seq, masks):                                                # :300-310
        baseline = self.vse(fc_feats, att_feats, seq, masks, True,
                            only_one_retrieval=self.only_one_retrieval)
        sc_loss = _sampleLogProbs * (retrieval_loss - baseline).detach().unsqueeze(1) * \
            _masks[:, 1:].detach().float()
        return baseline, sc_loss

    def no_baseline(self, retrieval_loss, _sampleLogProbs, _masks):                     # :312-319
        sc_loss = _sampleLogProbs * retrieval_loss.detach().unsqueeze(1) * \
            _masks[:, 1:].detach().float()
        return 0, sc_loss

    def loss_configuration(self, loss, sc_loss, baseline, retrieval_loss, _masks):      # :321-332
        sc_loss = sc_loss.sum() / _masks[:, 1:].float().sum()
        loss = loss + self.retrieval_reward_weight * sc_loss
        self._loss["retrieval_sc_loss"] = sc_loss.detach()
        self._loss["retrieval_loss"] = retrieval_loss.sum().detach()
        self._loss["retrieval_loss_greedy"] = baseline.sum().detach() \
            if torch.is_tensor(baseline) else baseline
        return loss

    def reinforce(self, fc_feats, att_feats, att_masks, seq, masks, data, loss):        # :334-341
        retrieval_loss, _seqs, _masks, gen_result, sample_logprobs, gen_masks = \
            self.reinforce_disc(fc_feats, att_feats, att_masks)
        return loss, gen_result, sample_logprobs, _masks, gen_result, gen_masks, _seqs, \
            retrieval_loss

    def st_and_ps_methods(self, fc_feats, att_feats, att_masks, seq, masks, data, loss):
        """gumbel / multinomial joint loss (:343-376), fused (see module docstring).  Returns
        (loss, word_index, logprobs, masks, _seqs) with the caption tensors left time-major and
        unsliced on the device (they only feed the out-of-scope CIDEr branch in the reference)."""
        spk, lis = self.caption_generator, self.vse
        # nothing differentiates the sampled ids' log-probabilities unless the CIDEr term is on:
        # the Gumbel pass may then keep z + G instead of z (no noise regenerated in backward)
        sp, st_mode = spk._sample_pass(att_feats, att_masks, 0, 1, 1,
                                       **({} if self.cider_optimization else {"store_perturbed": True}))  # :346-348
        assert st_mode
        B, V = sp.B, spk.vocab_size
        tok_sb = torch.cat([torch.full((1, B), V + 1, dtype=torch.int64, device=fc_feats.device),
                            sp.t["tok_out"][: sp.n_steps]], 0)                          # :360-370
        Pl = lis._params()
        is_ps = sp.ctx.mode in EN.PS_MODES
        emb16 = None
        if is_ps:
            # dense caption vectors: seqs_embed = _seqs @ embed.weight (VSEFCModel.py:102-104)
            emb16 = EN.caption_embed_dense(sp.t["soft16"], sp.n_steps, Pl, lis._packed.get(Pl), V + 1)
        lp = EN.listener_forward(Pl, lis._packed.get(Pl), fc_feats.detach().float().contiguous(),
                                 None if is_ps else tok_sb, sp.t["cap_len"], margin=lis.margin,
                                 only_one_retrieval=self.only_one_retrieval,
                                 no_imgnorm=bool(lis.img_enc.no_imgnorm), emb16=emb16,
                                 **lis._variant())                                      # :371-373
        if lis.keep_passes:
            lp.pinned = True
            lis._passes.append(lp)
        Ps = spk._params()
        logp = sp.t["logp"][: sp.n_steps]
        if torch.is_grad_enabled():
            if is_ps:
                loss_vse = _PsJointFn.apply(self, sp, lp, *_ordered_s(Ps),
                                            *[Pl[n] for n in EN.LISTENER_PARAM_NAMES])
            else:
                loss_vse, logp = _StJointFn.apply(self, sp, lp, *_ordered_s(Ps),
                                                  *[Pl[n] for n in EN.LISTENER_PARAM_NAMES])
        else:
            loss_vse = lp.t["loss"][0].clone()
        lis._loss["contrastive"] = loss_vse.detach()
        term = loss_vse * self.retrieval_reward_weight                                  # :374
        loss = term if (isinstance(loss, float) and loss == 0.0) else loss + term
        return loss, sp.t["tok_out"][: sp.n_steps], logp, sp.t["cap_len"], tok_sb

    # ------------------------------------------------------------------ CIDEr terms (:378-431)
    def gen_result_for_cider(self, fc_feats, att_feats, att_masks):
        """Sampled index captions with differentiable log-probabilities (:378-389), kept time-major
        [T, B] on the device (no host sync for the caption width)."""
        spk = self.caption_generator
        sp, _ = spk._sample_pass(att_feats, att_masks, 0, 1.0, 0)
        T = sp.n_steps
        if spk._needs_grad():
            logp = _SpeakerLossFn.apply(sp, sp.t["tok_fed"][1:T + 1].contiguous(), spk,
                                        *_ordered_s(spk._params()))
        else:
            logp = sp.t["logp"][:T]
        return sp.t["tok_out"][:T], logp

    def greedy_res_for_cider(self, fc_feats, att_feats, att_masks):
        """Greedy captions of the CURRENT mode of the module (the reference does not switch to eval
        here, :391-405, so dropout stays on in training), time-major [T, B]."""
        with torch.no_grad():
            sp, _ = self.caption_generator._sample_pass(att_feats, att_masks, 1, 1.0, 0)
        return sp.t["tok_out"][: sp.n_steps]

    def _time_major_ids(self, x):
        """[B, n] ids as returned by `sample` -> [T, B] (0-padded to the full length)."""
        T = self.caption_generator.seq_length
        x = x.detach().long()
        if x.size(1) < T:
            x = torch.nn.functional.pad(x, (0, T - x.size(1)))
        return x.t().contiguous()

    def traditional_cider(self, fc_feats, att_feats, att_masks, data, loss, gen_result, greedy_res,
                          sample_logprobs, gen_masks):
        """The reference's signature (:407-431): gen_result / greedy_res [B, n] ids and
        sample_logprobs [B, n] as returned by `sample`."""
        return self._cider_term(loss, data, fc_feats, self._time_major_ids(gen_result),
                                self._time_major_ids(greedy_res), sample_logprobs, False)

    def _cider_term(self, loss, data, fc_feats, hyp0, hyp1, logp, logp_time_major):
        """loss += cider_optimization * sum(logp * -reward * mask) / sum(mask)          (:407-431)
        with reward = CIDEr-D(sampled) - CIDEr-D(greedy) (or the sampled score alone when
        use_gen_cider_scores != 0) scored on the device; hyp0 / hyp1 int64 [T, B]."""
        if rewards.CiderD_scorer is None:
            rewards.init_scorer(getattr(self.opt, "cached_tokens", "corpus"))           # train.py:483
        B = fc_feats.size(0)
        gts = data.get("_coopcap_gts") if isinstance(data, dict) else None
        if gts is None:
            if not isinstance(data, dict) or "gts" not in data:
                raise ValueError("cider_optimization needs data['gts'] (dataloader.py:239)")
            gts = rewards.stage_gts(data["gts"], B, fc_feats.device)
        res = rewards.reward_on_device(rewards.CiderD_scorer, gts, hyp0.contiguous(), hyp1.contiguous(),
                                       differenced=(self.use_gen_cider_scores == 0))
        coef = res.coef if logp_time_major else res.coef[: logp.size(1)].t()
        loss_cider = (logp * coef).sum()
        self._loss["avg_reward"] = res.stats[0].float()
        self._loss["cider_greedy"] = res.stats[1].float()
        self._loss["loss_cider"] = loss_cider.detach()
        self._cider_last = res
        term = self.cider_optimization * loss_cider
        return term if (isinstance(loss, float) and loss == 0.0) else loss + term

    # ------------------------------------------------------------------ forward (:433-555)
    def forward(self, fc_feats, seq, masks, data, att_feats, att_masks, is_alternating=False,
                alternating_turn=None):
        if not is_alternating:
            gen = greedy_tb = None    # (ids [T,B], sampleLogprobs, time-major?) of the sampled captions
            # (:449-454) caption_loss_weight * loss_cap + vse_loss_weight * loss_vse; a term whose
            # weight is 0 is skipped instead of being materialised as a zero tensor, so the speaker
            # turn does not queue half a dozen one-element kernels between decode and listener
            loss = 0.0
            if self.caption_loss_weight > 0:
                loss = self.caption_loss_weight * self.ce_loss(fc_feats, att_feats, att_masks, seq, masks)
            if self.vse_loss_weight > 0:
                loss = loss + self.vse_loss_weight * self.vse_loss(
                    fc_feats, att_feats, seq, masks, only_one_retrieval=self.only_one_retrieval)
            if self.retrieval_reward_weight > 0:
                if self.retrieval_reward == "reinforce":
                    loss, gen_result, sample_logprobs, _masks, gen_result, gen_masks, _seqs, \
                        retrieval_loss = self.reinforce(fc_feats, att_feats, att_masks, seq, masks,
                                                        data, loss)
                    if self.reinforce_baseline_type == "greedy":
                        baseline, sc_loss, greedy_res = self.greedy_baseline(
                            fc_feats, att_feats, att_masks, retrieval_loss, _seqs, sample_logprobs,
                            _masks)
                        greedy_tb = self._time_major_ids(greedy_res)
                    elif self.reinforce_baseline_type == "gt":
                        baseline, sc_loss = self.gt_baseline(
                            fc_feats, att_feats, att_masks, retrieval_loss, _seqs, sample_logprobs,
                            _masks, seq, masks)
                    else:
                        baseline, sc_loss = self.no_baseline(retrieval_loss, sample_logprobs, _masks)
                    loss = self.loss_configuration(loss, sc_loss, baseline, retrieval_loss, _masks)
                    gen = (self._time_major_ids(gen_result), sample_logprobs, False)
                else:
                    loss, tok_tb, logp_tb, *_ = self.st_and_ps_methods(
                        fc_feats, att_feats, att_masks, seq, masks, data, loss)
                    if self.retrieval_reward in ("gumbel", "multinomial"):
                        gen = (tok_tb, logp_tb, True)
            if self.cider_optimization:                                                 # :490-503
                if gen is None:    # no sampled captions yet, or a partial-sampling mode (:492-496)
                    gen = self.gen_result_for_cider(fc_feats, att_feats, att_masks) + (True,)
                if greedy_tb is None:
                    greedy_tb = self.greedy_res_for_cider(fc_feats, att_feats, att_masks)
                loss = self._cider_term(loss, data, fc_feats, gen[0], greedy_tb, gen[1], gen[2])
            return loss if torch.is_tensor(loss) else self._zero(fc_feats)
        if alternating_turn == "speaker":                                               # :508-526
            if self.retrieval_reward == "reinforce":
                self.changeModelUpdateStatus({"vseModel": False, "captionModel": True})
            old = self.getLossFlags()
            self.setLossFlages(VSEWeight=0, MLEWeight=old[1], ciderFlag=old[2], DISCWeight=old[3])
            try:
                return self.forward(fc_feats, seq, masks, data, att_feats, att_masks,
                                    is_alternating=False, alternating_turn=alternating_turn)
            finally:
                self.setLossFlages(*old)
        if alternating_turn == "listener":                                              # :528-555
            self.changeModelUpdateStatus({"vseModel": True, "captionModel": False})
            old = self.getLossFlags()
            self.setLossFlages(VSEWeight=old[0], MLEWeight=0, ciderFlag=0, DISCWeight=0)
            try:
                with torch.no_grad():
                    _seqs, _ = self.caption_generator.sample(
                        fc_feats, att_feats, att_masks, {"sample_max": 0, "temperature": 1})
                _masks = self._caption_masks(_seqs)
                _seqs = self._with_bos(_seqs)
                return self.forward(fc_feats, _seqs, _masks, data, att_feats, att_masks,
                                    is_alternating=False, alternating_turn=alternating_turn)
            finally:
                self.setLossFlages(*old)
        raise ValueError(f"alternating_turn must be 'speaker' or 'listener', got {alternating_turn!r}")

    def sample(self, fc_feats, att_feats, att_masks, opt={}):                           # :557-560
        return self.caption_generator.sample(fc_feats, att_feats, att_masks, opt)

    def loss(self):                                                                     # :562-568
        out = {}
        out.update(self._loss)
        out.update({"cap_" + k: v for k, v in self.caption_generator._loss.items()})
        out.update({"vse_" + k: v for k, v in self.vse._loss.items()})
        return out

    def changeModelUpdateStatus(self, gradientsDic, printWeights=False):
        """requires_grad toggling of :571-633; the deep-copy consistency check (:584-586,
        :680-685) is not reproduced."""
        for p in self.vse.parameters():
            p.requires_grad = bool(gradientsDic["vseModel"])
        for p in self.caption_generator.parameters():
            p.requires_grad = bool(gradientsDic["captionModel"])
