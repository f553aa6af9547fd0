"""Drop-in listener: VSEFCModel with the reference's interface (models/VSEFCModel.py).

Same constructor, parameter names / shapes (`img_enc.fc`, `txt_enc.embed`, `txt_enc.rnn.*_l0`),
same `forward(fc_feats, att_feats, seq, masks, whole_batch=False, only_one_retrieval='off')`
(VSEFCModel.py:230-241).  `seq` is int64 [B, S] ids or a float one-hot tensor [B, S, V+2]
(VSEFCModel.py:102-106); a dense input (one-hot or soft vectors) is embedded by the dense
contraction `seq @ embed.weight` like the reference does, and its gradient is returned densely
(demb . W_emb^T) so foreign callers still see the reference's autograd graph.
All arithmetic runs in libcoopcap; there is no CPU path.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn as nn

from .. import engine as EN
from .. import ops


class EncoderImage(nn.Module):
    """VSEFCModel.py:20-54: fc -> l2norm (-> abs)."""

    def __init__(self, opt):
        super().__init__()
        self.embed_size = opt.vse_embed_size
        self.no_imgnorm = opt.vse_no_imgnorm
        self.use_abs = opt.vse_use_abs
        self.fc_feat_size = opt.fc_feat_size
        self.fc = nn.Linear(self.fc_feat_size, self.embed_size)
        r = (6.0 ** 0.5) / ((self.fc.in_features + self.fc.out_features) ** 0.5)   # :32-38
        self.fc.weight.data.uniform_(-r, r)
        self.fc.bias.data.fill_(0)


class EncoderText(nn.Module):
    """VSEFCModel.py:57-81: embedding + 1-layer GRU (parameter holder)."""

    def __init__(self, opt):
        super().__init__()
        self.use_abs = opt.vse_use_abs
        self.input_encoding_size = opt.input_encoding_size
        self.embed_size = opt.vse_embed_size
        self.num_layers = opt.vse_num_layers
        self.rnn_type = opt.vse_rnn_type
        self.vocab_size = opt.vocab_size
        self.pool_type = getattr(opt, "vse_pool_type", "")
        self.embed = nn.Embedding(self.vocab_size + 2, self.input_encoding_size)
        self.rnn = getattr(nn, self.rnn_type.upper())(self.input_encoding_size, self.embed_size,
                                                      self.num_layers, batch_first=True)
        self.embed.weight.data.uniform_(-0.1, 0.1)                                # :80-81


class ContrastiveLoss(nn.Module):
    """VSEFCModel.py:149-165 (hyper-parameter holder; the loss runs in csrc/listener.cu)."""

    def __init__(self, opt):
        super().__init__()
        self.margin = opt.vse_margin
        self.measure = opt.vse_measure
        self.max_violation = opt.vse_max_violation


def _ordered(P):
    return [P[n] for n in EN.LISTENER_PARAM_NAMES]


class _ListenerFn(torch.autograd.Function):
    """Listener loss (scalar [ ] or per-row [B]) as a differentiable function of the listener
    parameters and, optionally, of a dense one-hot caption tensor."""

    @staticmethod
    def forward(ctx, owner, lp, whole_batch, dense_seq, *params):
        ctx.owner, ctx.lp, ctx.whole_batch = owner, EN.retain(lp), whole_batch
        ctx.dense = dense_seq is not None and dense_seq.requires_grad
        ctx.dense_shape = None if dense_seq is None else dense_seq.shape
        out = lp.t["loss_rows"] if whole_batch else lp.t["loss"][0]
        return out.clone()

    @staticmethod
    def backward(ctx, g):
        owner, lp = ctx.owner, ctx.lp
        P = owner._params()
        need = any(p.requires_grad for p in P.values())
        g = g.contiguous().float()
        kw = dict(g_rows=g) if ctx.whole_batch else dict(g_loss=g.reshape(1))
        G, demb16 = EN.listener_backward(lp, P, need_param_grads=need, **kw)
        g_seq = None
        if ctx.dense:
            B, S, V2 = ctx.dense_shape
            w16 = owner._packed.get(P)["w_emb16"]
            flat = torch.empty(S * B, V2, device=g.device)
            if V2 % 4:   # TMA needs a 16-byte row pitch: compute into a padded buffer
                pad = torch.empty(S * B, (V2 + 3) // 4 * 4, device=g.device)
                ops.gemm(demb16.view(S * B, -1), w16, S * B, V2, demb16.shape[-1], out=pad)
                flat = pad[:, :V2]
            else:
                ops.gemm(demb16.view(S * B, -1), w16, S * B, V2, demb16.shape[-1], out=flat)
            g_seq = flat.view(S, B, V2).transpose(0, 1).contiguous()
        x16 = getattr(lp, "dense_x16", None)
        if need and x16 is not None:
            # embedding weight gradient of the dense contraction: x^T . demb   ([V2, S*B] x [S*B, E])
            S, B, V2p = x16.shape
            V2 = P["txt_enc.embed.weight"].shape[0]
            ops.gemm(x16.view(S * B, V2p), demb16.view(S * B, -1), V2, demb16.shape[-1], S * B,
                     a_major=1, b_major=1, out=G["txt_enc.embed.weight"])
        grads = tuple((G[n].view_as(P[n]) if need else None) for n in EN.LISTENER_PARAM_NAMES)
        EN.release(lp)
        ctx.lp = None
        return (None, None, None, g_seq) + grads


class VSEFCModel(nn.Module):
    def __init__(self, opt):
        super().__init__()
        self.loss_type = getattr(opt, "vse_loss_type", "contrastive")
        if self.loss_type != "contrastive":
            raise NotImplementedError("only the contrastive listener loss is on the hot path")
        if opt.vse_rnn_type.lower() != "gru" or opt.vse_num_layers != 1:
            raise NotImplementedError("the listener is a 1-layer GRU (run_joint.sh defaults)")
        self.pool_type = getattr(opt, "vse_pool_type", "last") or "last"          # :70,:115-127
        if self.pool_type not in EN.POOL:
            raise ValueError(f"unknown vse_pool_type {self.pool_type!r}")
        self.max_violation = bool(getattr(opt, "vse_max_violation", 1))           # :163,:190-193
        self.use_abs = bool(getattr(opt, "vse_use_abs", 0))                        # :31,:69
        self.img_enc = EncoderImage(opt)
        self.txt_enc = EncoderText(opt)
        self.contrastive_loss = ContrastiveLoss(opt)
        self.margin = opt.vse_margin
        self.embed_size = opt.vse_embed_size
        self._loss = {}
        self.keep_passes = False          # tests: retain every ListenerPass in self._passes
        self._passes = []
        self._packed = EN.PackedListener()

    def _params(self) -> Dict[str, torch.Tensor]:
        r = self.txt_enc.rnn
        return {
            "img_enc.fc.weight": self.img_enc.fc.weight, "img_enc.fc.bias": self.img_enc.fc.bias,
            "txt_enc.embed.weight": self.txt_enc.embed.weight,
            "txt_enc.rnn.weight_ih_l0": r.weight_ih_l0, "txt_enc.rnn.weight_hh_l0": r.weight_hh_l0,
            "txt_enc.rnn.bias_ih_l0": r.bias_ih_l0, "txt_enc.rnn.bias_hh_l0": r.bias_hh_l0,
        }

    def _variant(self) -> dict:
        return dict(pool_type=self.pool_type, use_abs=self.use_abs, max_violation=self.max_violation)

    def _forward_ids(self, fc_feats, tok_sb, lens, whole_batch, only_one_retrieval, dense_seq=None):
        """tok_sb int64 [S, B] time-major, lens int32 [B]; or dense_seq float [B, S, V+2]."""
        if not fc_feats.is_cuda:
            raise EN._lib.CoopcapError("VSEFCModel runs on CUDA only (no CPU path)")
        P = self._params()
        packed = self._packed.get(P)
        emb16 = x16 = None
        if dense_seq is not None:
            # seqs.dim() > 2: seqs_embed = seqs @ embed.weight (VSEFCModel.py:102-104); the dense
            # tensor is re-laid time-major in bf16 (row pitch padded to 16 bytes for TMA)
            B, S, V2 = dense_seq.shape
            d = packed["dims"]
            if V2 != d.V2:
                raise EN._lib.CoopcapError(f"dense captions have width {V2}, the embedding has {d.V2} rows")
            x16 = torch.zeros(S, B, (V2 + 7) // 8 * 8, dtype=torch.bfloat16, device=dense_seq.device)
            x16[:, :, :V2] = dense_seq.detach().transpose(0, 1)
            emb16 = torch.empty(S, B, d.E, dtype=torch.bfloat16, device=dense_seq.device)
            ops.gemm(x16.view(S * B, -1), packed["w_emb16"], S * B, d.E, V2, b_major=1,
                     out16=emb16.view(S * B, d.E))
        lp = EN.listener_forward(P, packed, fc_feats.detach().float().contiguous(), tok_sb, lens,
                                 margin=self.margin, only_one_retrieval=only_one_retrieval,
                                 no_imgnorm=bool(self.img_enc.no_imgnorm), emb16=emb16,
                                 **self._variant())
        lp.dense_x16 = x16
        if self.keep_passes:
            lp.pinned = True
            self._passes.append(lp)
        needs = torch.is_grad_enabled() and (
            any(p.requires_grad for p in P.values()) or
            (dense_seq is not None and dense_seq.requires_grad))
        if needs:
            out = _ListenerFn.apply(self, lp, whole_batch, dense_seq, *_ordered(P))
        else:
            out = (lp.t["loss_rows"] if whole_batch else lp.t["loss"][0]).clone()
        return out, lp

    def forward(self, fc_feats, att_feats, seq, masks, whole_batch=False,
                only_one_retrieval="off"):
        """VSEFCModel.py:230-241 (att_feats is ignored, as in the reference)."""
        lens = (masks > 0).sum(1).to(torch.int32).contiguous()                     # :84
        dense, tok_sb = None, None
        if seq.dim() > 2:
            dense = seq
        else:
            tok_sb = seq.long().t().contiguous()
        loss, _ = self._forward_ids(fc_feats, tok_sb, lens, whole_batch, only_one_retrieval, dense)
        if not whole_batch:
            self._loss["contrastive"] = loss.detach()                               # :238-239
        return loss
