"""CPU / gloo, world_size 2: the data-parallel host logic (flat gradient bucket, one all-reduce,
1/N scaling) of optimizer.FlatAdam -- SURVEY.md §8(e)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cooperativeimagecaptioning_b200 import optimizer as OPT
        from cooperativeimagecaptioning_b200._lib import CoopcapError
        torch.manual_seed(0)                                  # identical weights on every rank
        net = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3))
        w0 = [p.detach().clone() for p in net.parameters()]
        opt = OPT.FlatAdam(net.parameters(), lr=1e-3, grad_clip=0.1)
        # parameters are views of the flat bucket; gradients join theirs in gather_grads()
        for p, o in zip(opt.params, opt.offsets):
            assert p.data_ptr() == opt.flat_param.data_ptr() + 4 * o
            assert p.grad is None
            assert o % 4 == 0
        for p, w in zip(net.parameters(), w0):
            assert torch.equal(p, w)
        # rank-local "shard" loss -> rank-local gradients, gathered into the bucket
        opt.zero_grad()
        g = torch.Generator().manual_seed(100 + rank)
        x = torch.randn(4, 7, generator=g)
        net(x).square().sum().backward()
        want = [p.grad.clone() for p in net.parameters()]
        opt.gather_grads()
        for p, o, w in zip(opt.params, opt.offsets, want):
            assert p.grad.data_ptr() == opt.flat_grad.data_ptr() + 4 * o and torch.equal(p.grad, w)
        # a parameter without a gradient contributes zeros, a second backward accumulates
        opt.zero_grad()
        net[1](torch.randn(2, 5, generator=g)).sum().backward()
        net[1](torch.randn(2, 5, generator=g)).sum().backward()
        opt.gather_grads()
        assert float(net[0].weight.grad.abs().sum()) == 0.0 and float(net[1].bias.grad[0]) == 4.0
        opt.zero_grad()
        net(x).square().sum().backward()
        opt.gather_grads()
        local = opt.flat_grad.clone()
        n = opt.all_reduce()
        assert n == world
        gathered = [torch.zeros_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        assert torch.allclose(opt.flat_grad, sum(gathered))   # one SUM all-reduce of the bucket
        # the fused clamp+Adam kernel is CUDA-only: the CPU path must fail loudly
        try:
            opt.step(world_size=n)
            raise AssertionError("FlatAdam.step must not run on CPU")
        except CoopcapError:
            pass
        # checkpoint layout is torch.optim-shaped
        sd = opt.state_dict()
        assert set(sd) == {"state", "param_groups"} and len(sd["state"]) == 0   # no step yet
        opt.step_count = 1
        assert len(opt.state_dict()["state"]) == 4
        # a parameter shared by two optimizers (--share_embed): the second owner copies the
        # gradient the first one already summed over the ranks and must not sum it again
        shared = torch.nn.Parameter(torch.zeros(6))
        a_own, b_own = torch.nn.Parameter(torch.zeros(3)), torch.nn.Parameter(torch.zeros(5))
        A = OPT.FlatAdam([shared, a_own], lr=1e-3)
        Bo = OPT.FlatAdam([b_own, shared], lr=1e-3)
        assert Bo.foreign == [False, True]
        for reduce_a_first in (True, False):
            A.zero_grad()
            Bo.zero_grad()
            ((shared.sum() + a_own.sum() + b_own.sum()) * float(rank + 1)).backward()
            if reduce_a_first:                      # gumbel: both agents step (optimizer.py:233-237)
                A.all_reduce()
            Bo.all_reduce()                         # reinforce listener turn: only this one steps
            total = float(sum(r + 1 for r in range(world)))
            assert torch.equal(Bo.flat_grad[Bo.offsets[1]:Bo.offsets[1] + 6], torch.full((6,), total))
            assert torch.equal(Bo.flat_grad[:5], torch.full((5,), total))
            if reduce_a_first:
                assert torch.equal(A.flat_grad[:6], torch.full((6,), total))
        out.put((rank, float(opt.flat_grad.abs().sum())))
    finally:
        dist.destroy_process_group()


def test_flat_bucket_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    res = dict(q.get() for _ in range(2))
    assert abs(res[0] - res[1]) < 1e-6                        # both ranks hold the same sum
