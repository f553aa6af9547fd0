"""GPU: one straight-through joint step (speaker sample -> listener loss -> backward through both)
driven through the C ABI, against the CPU oracle with identical injected noise (replay mode)."""
import pytest
import torch

from oracle import joint as OJ
from oracle import speaker as OS
from oracle import synth
from gpu_util import REAL, branch_replay, check_near_ties, cuda_params, pack_keep, u8

pytestmark = pytest.mark.gpu

BF16_TOL = 2e-2   # north star: 2e-2 relative (bf16 operands, fp32 accumulation)
NEAR_TIE = 2e-2   # maxout / ReLU decisions may differ only below 2% of the median margin


def _grad_err(g, ref):
    g, ref = g.detach().double().cpu(), ref.detach().double()
    return float((g - ref).abs().max()) / max(float(ref.abs().max()), 1e-30)


@pytest.mark.parametrize("mode,varlen,dropout,tau,eos_bias", [
    ("gumbel", False, False, 1.0, 0.0),
    ("gumbel", True, True, 0.75, 7.5),
    ("multinomial", True, True, 1.0, 7.5),
])
def test_st_joint_step_matches_oracle(mode, varlen, dropout, tau, eos_bias):
    from cooperativeimagecaptioning_b200 import engine as EN
    d = REAL
    B, L, seed, T, V = 10, 7, 21, d.seq_length, d.vocab_size
    Ps = synth.speaker_params(d, seed=seed, eos_bias=eos_bias)
    Pl = synth.listener_params(d, seed=seed + 1)
    batch = synth.make_batch(d, B, L, seed + 2, varlen=varlen, min_regions=2)
    noise = synth.make_noise(d, B, L, seed + 3, dropout=dropout, gumbel=(mode == "gumbel"),
                             multinomial=(mode == "multinomial"))
    drop_p = 0.5 if dropout else 0.0
    cfg = OJ.JointCfg(drop_p=drop_p, retrieval_reward=mode, gumbel_temp=tau, multinomial_temp=tau,
                      retrieval_reward_weight=0.01)
    # pass 1: the oracle's own draws (all steps), pass 2: replay with the reference's early stop
    free = OS.sample(Ps, batch.att_feats, batch.att_masks, mode=mode, seq_length=T, vocab_size=V,
                     noise=noise, drop_p=drop_p, sample_max=0, use_one_hot=1, gumbel_temp=tau,
                     multinomial_temp=tau, keep_all_steps=True)
    forced_bt = torch.stack(free.tokens_raw, 1)
    # CUDA path
    Pc, Plc = cuda_params(Ps), cuda_params(Pl)
    packed_s, packed_l = EN.PackedSpeaker().get(Pc), EN.PackedListener().get(Plc)
    masks_c = None if batch.att_masks is None else batch.att_masks.cuda()
    off, NL = EN.region_offsets(masks_c, B, L)
    rnd = EN.SpeakerRandom(seed=1, drop_p=drop_p)
    if dropout:
        rnd.keep_att = pack_keep(noise.drop_att, batch.att_masks)
        rnd.keep_embed = u8(noise.drop_embed)
        rnd.keep_core = u8(noise.drop_core)
    rnd.noise = (noise.U if mode == "gumbel" else noise.E).cuda().contiguous()
    cmode = EN.MODE_ST_GUMBEL if mode == "gumbel" else EN.MODE_ST_MULTINOMIAL
    sp = EN.speaker_forward(Pc, packed_s, batch.att_feats.cuda(), off, NL, n_steps=T, mode=cmode,
                            inv_tau=1.0 / tau, start_token=V + 1, rnd=rnd,
                            forced=forced_bt.t().contiguous().cuda())
    tok_sb = torch.cat([torch.full((1, B), V + 1, dtype=torch.int64, device="cuda"),
                        sp.t["tok_out"]], 0).contiguous()
    lp = EN.listener_forward(Plc, packed_l, batch.fc_feats.cuda(), tok_sb, sp.t["cap_len"])
    rn = branch_replay(sp, batch.att_masks, noise)
    Pso = {k: v.clone().requires_grad_(True) for k, v in Ps.items()}
    Plo = {k: v.clone().requires_grad_(True) for k, v in Pl.items()}
    loss, res, masks, loss_vse = OJ.st_joint_loss(Pso, Plo, batch.fc_feats, batch.att_feats,
                                                  batch.att_masks, rn, cfg, forced_bt)
    gs = torch.autograd.grad(loss, list(Pso.values()) + list(Plo.values()), allow_unused=True)
    g_ref_s = {k: (torch.zeros_like(v) if g is None else g)
               for (k, v), g in zip(Pso.items(), gs[:len(Pso)])}
    g_ref_l = {k: (torch.zeros_like(v) if g is None else g)
               for (k, v), g in zip(Plo.items(), gs[len(Pso):])}

    print(check_near_ties(rn, batch.att_masks, NEAR_TIE))
    n = int(sp.t["n_out"].item())
    assert n == res.seq.shape[1]
    assert torch.equal(sp.t["tok_out"][:n].t().cpu(), res.seq)
    assert torch.equal(sp.t["cap_len"].cpu().long(), (masks > 0).sum(1))
    got = float(lp.t["loss"].item())
    assert abs(got - float(loss_vse)) <= BF16_TOL * abs(float(loss_vse)), (got, float(loss_vse))

    w = torch.full((1,), cfg.retrieval_reward_weight, device="cuda")
    Gl, demb16 = EN.listener_backward(lp, Plc, g_loss=w)
    dz16 = EN.st_logit_grads(sp, demb16[1:T + 1].contiguous(), packed_l["w_emb16"])
    Gs = EN.speaker_backward(sp, dz16, Pc)
    torch.cuda.synchronize()
    worst = 0.0
    for k, ref in list(g_ref_l.items()) + list(g_ref_s.items()):
        g = (Gl if k in Gl else Gs)[k]
        if float(ref.abs().max()) == 0.0:
            assert float(g.abs().max()) <= 1e-9, k
            continue
        if k.endswith("alpha_net.bias"):
            assert float(g.abs().max()) <= 1e-7
            continue
        e = _grad_err(g, ref)
        gd, rd = g.detach().double().cpu().flatten(), ref.detach().double().flatten()
        l2 = float((gd - rd).norm() / rd.norm())
        cos = float((gd @ rd) / (gd.norm() * rd.norm()))
        print(f"  {mode} {k:45s} rel err {e:.3e} l2 {l2:.3e} cos {cos:.6f} ratio {float(gd.norm()/rd.norm()):.4f} (|ref|max {float(ref.abs().max()):.3e})")
        worst = max(worst, l2)
    # alpha_net.bias: analytically 0 (softmax shift invariance); the reference gets rounding noise
    assert worst <= BF16_TOL, worst
