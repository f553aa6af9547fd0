"""GPU, 2 ranks over NCCL: data-parallel parity as SURVEY.md §8(e) defines it -- after the flat
gradient all-reduce, `bucket / N` equals the MEAN of the per-shard single-process oracle gradients
(each shard with its own injected noise, batch-local negatives), and the `÷N → clamp → Adam` step
that follows moves every rank's parameters identically.

Needs two visible GPUs (`gpurun --gpus 2 -- python -m pytest tests/test_gpu_dist.py -m gpu`); on
a one-GPU box the test skips.  The record is written to gpurun_out/parity/dp2_nccl.json."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu

ROWS, REGIONS, SEED = 12, 8, 401


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _shard(rank):
    """Everything one rank's single-process run needs, built identically on every rank."""
    from oracle import synth
    d = synth.Dims()
    batch = synth.make_batch(d, ROWS, REGIONS, SEED + 10 * rank + 2, varlen=True, min_regions=2)
    noise = synth.make_noise(d, ROWS, REGIONS, SEED + 10 * rank + 3, dropout=True, gumbel=True)
    return batch, noise


def _oracle_shard_grads(Ps, Pl, batch, noise, forced, replay, hinge):
    from oracle import joint as OJ
    cfg = OJ.JointCfg(drop_p=0.5, retrieval_reward="gumbel", gumbel_temp=0.75)
    Pso = {k: v.clone().requires_grad_(True) for k, v in Ps.items()}
    Plo = {k: v.clone().requires_grad_(True) for k, v in Pl.items()}
    loss, _, _, _ = OJ.st_joint_loss(Pso, Plo, batch.fc_feats, batch.att_feats, batch.att_masks,
                                     replay if replay is not None else noise, cfg, forced,
                                     hinge_replay=hinge)
    ts = list(Pso.values()) + list(Plo.values())
    names = ["caption_generator." + k for k in Pso] + ["vse." + k for k in Plo]
    gs = torch.autograd.grad(loss, ts, allow_unused=True)
    return float(loss), {n: (torch.zeros_like(t) if g is None else g) for n, t, g in zip(names, ts, gs)}


def _worker(rank, world, port, out):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.dirname(here), here):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from oracle import speaker as OS
        from oracle import synth
        from oracle.ref_loader import reference_opt
        import cooperativeimagecaptioning_b200.models as models
        from cooperativeimagecaptioning_b200 import engine as EN
        from cooperativeimagecaptioning_b200 import optimizer as OPT
        from gpu_util import branch_replay, grad_report, hinge_replay_of, pack_keep, u8
        d = synth.Dims()
        Ps = synth.speaker_params(d, seed=SEED, eos_bias=7.5)
        Pl = synth.listener_params(d, seed=SEED + 1)
        opt = reference_opt(retrieval_reward="gumbel", gumbel_temp=0.75, drop_prob_lm=0.5,
                            batch_size=ROWS, learning_rate=5e-4, grad_clip=0.1)
        model = models.AlternatingJointModel(opt)
        sd = {"caption_generator." + k: v for k, v in Ps.items()}
        sd.update({"vse." + k: v for k, v in Pl.items()})
        model.load_state_dict(sd)
        model.cuda().train()
        optim = OPT.define_optimizer(model, opt)
        batch, noise = _shard(rank)
        with torch.no_grad():
            free = OS.sample(Ps, batch.att_feats, batch.att_masks, mode="gumbel", seq_length=d.seq_length,
                             vocab_size=d.vocab_size, noise=noise, drop_p=0.5, sample_max=0,
                             use_one_hot=1, gumbel_temp=0.75, keep_all_steps=True)
        forced = torch.stack(free.tokens_raw, 1)
        spk = model.caption_generator
        spk.injected = EN.SpeakerRandom(seed=1, drop_p=0.5, keep_att=pack_keep(noise.drop_att, batch.att_masks),
                                        keep_embed=u8(noise.drop_embed), keep_core=u8(noise.drop_core),
                                        noise=noise.U.cuda().contiguous())
        spk.forced_tokens = forced.cuda()
        spk.keep_passes = model.vse.keep_passes = True
        optim.zero_grad()
        loss = model(batch.fc_feats.cuda(), batch.labels.cuda(), batch.masks.cuda(), None,
                     batch.att_feats.cuda(), batch.att_masks.cuda(), is_alternating=True,
                     alternating_turn="speaker")
        loss.backward()
        # this rank's oracle gradient (decisions of its own CUDA pass replayed), shared with the peer
        rn = branch_replay(spk._passes[0], batch.att_masks, noise)
        hr = hinge_replay_of(model.vse._passes[0])
        loss_ref, mine = _oracle_shard_grads(Ps, Pl, batch, noise, forced, rn, hr)
        names = sorted(mine)
        flat = torch.cat([mine[n].flatten() for n in names]).cuda()
        both = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(both, flat)
        mean_ref = (sum(both) / world).cpu()
        n_ranks = optim.all_reduce()                       # the path's one collective
        assert n_ranks == world
        got = {n: (p.grad / world).clone() for n, p in model.named_parameters()}
        ref, o = {}, 0
        for n in names:
            k = mine[n].numel()
            ref[n] = mean_ref[o:o + k].view_as(mine[n])
            o += k
        rep = grad_report(got, ref)
        worst = max(rep.items(), key=lambda kv: kv[1]["l2"])
        assert abs(float(loss) - loss_ref) <= 2e-2 * abs(loss_ref)
        assert worst[1]["l2"] <= 2e-2 and min(v["cos"] for v in rep.values()) >= 0.9995, worst
        # ÷N -> clamp -> Adam: every rank ends with bit-identical parameters
        optim.step(world_size=n_ranks)
        digest = torch.cat([p.detach().flatten()[:64] for p in model.parameters()]).double().sum().reshape(1)
        peers = [torch.empty_like(digest) for _ in range(world)]
        dist.all_gather(peers, digest)
        assert all(torch.equal(peers[0], q) for q in peers)
        # ... and they are the reference's clamp + Adam step applied to the averaged bucket
        moved = 0.0
        for n, p in model.named_parameters():
            g = got[n].clamp(-0.1, 0.1)
            m, v = 0.1 * g, 0.001 * g * g
            want = sd[n].cuda() - (5e-4 / 0.1) * m / (v.sqrt() / (0.001 ** 0.5) + 1e-8)
            moved = max(moved, float((p.detach() - want).abs().max()))
        assert moved <= 1e-6, moved
        out.put((rank, dict(worst_tensor=worst[0], worst_l2=worst[1]["l2"],
                            worst_cos=min(v["cos"] for v in rep.values()), loss=float(loss),
                            loss_ref=loss_ref, max_param_dev_after_step=moved)))
    finally:
        dist.destroy_process_group()


def test_dp_gradient_is_the_mean_of_the_shard_oracle_gradients():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    import torch.multiprocessing as mp
    from gpu_util import write_report
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    res = dict(q.get() for _ in range(2))
    write_report("dp2_nccl", {f"rank{r}": v for r, v in res.items()})
    assert abs(res[0]["worst_l2"] - res[1]["worst_l2"]) < 1e-9     # both ranks hold the same bucket
