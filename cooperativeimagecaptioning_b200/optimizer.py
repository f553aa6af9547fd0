"""Optimizer policy of the reference (optimizer.py:25-27,224-242 + misc/utils.py:65-69) as one
fused pass per agent, plus the data-parallel gradient exchange.

`FlatAdam` is a torch.optim.Optimizer-shaped object (`zero_grad`, `step`, `state_dict`,
`load_state_dict`, `param_groups[0]['lr']` for misc/utils.set_lr) whose parameters, gradients and
Adam moments live in flat fp32 buckets:

    all-reduce(sum) over ranks  ->  g /= world_size  ->  g = clamp(g, +-grad_clip)  ->  Adam

The all-reduce is the path's only collective (torch.distributed / NCCL over NVLink); the rest is
one hand-written kernel over the bucket (coopcap_clamp_adam).  Order matters: the clamp is
non-linear, so it runs after the average -- N ranks reproduce the mean of N single-process
reference gradients, then the reference's clamp + Adam (optimizer.py:237-241).

`define_optimizer(model, opt)` / `zeroing_optimizer` / `update_optimizer` keep the reference's
call shapes so train.py's loop reads the same.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional

import torch

from . import engine as EN


class FlatAdam:
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float, weight_decay: float = 0.0,
                 betas=(0.9, 0.999), eps: float = 1e-8, grad_clip: float = 0.0,
                 process_group=None):
        self.params: List[torch.nn.Parameter] = [p for p in params]
        if not self.params:
            raise ValueError("FlatAdam: no parameters")
        dev = self.params[0].device
        # the bucket plumbing (flattening, all-reduce) is device-agnostic torch code and is
        # exercised on CPU/gloo by tests/test_dist_cpu.py; step() itself needs the CUDA kernel
        # 16-byte aligned segments so every parameter view is vector-load friendly
        self.offsets, n = [], 0
        for p in self.params:
            self.offsets.append(n)
            n += (p.numel() + 3) // 4 * 4
        self.numel = n
        self.flat_param = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self._grad_views = []
        for p, o in zip(self.params, self.offsets):
            view = self.flat_param[o:o + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view                                    # parameters become views of the bucket
            self._grad_views.append(self.flat_grad[o:o + p.numel()].view_as(p))
            p.grad = None
        self.param_groups = [dict(lr=lr, weight_decay=weight_decay, betas=betas, eps=eps,
                                  params=self.params)]
        self.grad_clip = grad_clip
        self.step_count = 0
        self.process_group = process_group

    # reference call shape: optimizer.zero_grad()
    def zero_grad(self, set_to_none: bool = True):
        """Gradients are dropped, not zero-filled: the first gradient a parameter receives in the
        next backward pass is then adopted by autograd as `p.grad` without an accumulation kernel
        (24 read-modify-write passes over the 104 MB bucket per step otherwise); `gather_grads`
        moves them into the flat bucket with one multi-tensor copy."""
        for p in self.params:
            p.grad = None

    def gather_grads(self):
        """p.grad of every parameter -> its segment of the flat gradient bucket (zeros where a
        parameter received no gradient); afterwards p.grad IS that segment."""
        src, dst = [], []
        for p, v in zip(self.params, self._grad_views):
            g = p.grad
            if g is None:
                v.zero_()
            elif g.data_ptr() != v.data_ptr():
                src.append(g.detach().to(torch.float32).reshape(v.shape))
                dst.append(v)
            p.grad = v
        if src:
            torch._foreach_copy_(dst, src)

    def all_reduce(self):
        """Sum the gradient bucket over the data-parallel ranks (no-op for a single process)."""
        import torch.distributed as dist
        self.gather_grads()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.process_group) > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.process_group)
            return dist.get_world_size(self.process_group)
        return 1

    def step(self, grad_clip: Optional[float] = None, world_size: Optional[int] = None):
        """clamp + Adam; `world_size` > 1 divides the (already all-reduced) gradient first."""
        if self.flat_param.device.type != "cuda":
            raise EN._lib.CoopcapError("FlatAdam.step needs CUDA parameters (there is no CPU path)")
        if world_size is None:
            world_size = self.all_reduce()
        else:
            self.gather_grads()
        g = self.param_groups[0]
        self.step_count += 1
        clip = self.grad_clip if grad_clip is None else grad_clip
        EN.clamp_adam_(self.flat_param, self.flat_grad, self.exp_avg, self.exp_avg_sq,
                       step=self.step_count, lr=g["lr"], grad_scale=1.0 / world_size, clip=clip,
                       beta1=g["betas"][0], beta2=g["betas"][1], eps=g["eps"],
                       weight_decay=g["weight_decay"])
        # the kernel wrote the parameters through raw pointers: tell the packed bf16 operand
        # caches (engine.PackedSpeaker / PackedListener) that the masters changed
        EN.bump_weights_epoch()

    # torch.optim-compatible checkpoint layout (optimizer.py:191-221 saves optimizer.state_dict())
    def state_dict(self) -> Dict:
        state = {}
        for i, (p, o) in enumerate(zip(self.params, self.offsets)):
            n = p.numel()
            state[i] = dict(step=torch.tensor(float(self.step_count)),
                            exp_avg=self.exp_avg[o:o + n].view_as(p).clone(),
                            exp_avg_sq=self.exp_avg_sq[o:o + n].view_as(p).clone())
        g = self.param_groups[0]
        return dict(state=state, param_groups=[dict(
            lr=g["lr"], betas=g["betas"], eps=g["eps"], weight_decay=g["weight_decay"],
            amsgrad=False, params=list(range(len(self.params))))])

    def load_state_dict(self, sd: Dict):
        for i, (p, o) in enumerate(zip(self.params, self.offsets)):
            st = sd["state"].get(i)
            if st is None:
                continue
            n = p.numel()
            self.exp_avg[o:o + n].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
            self.step_count = int(float(st["step"]))
        g = sd["param_groups"][0]
        self.param_groups[0].update(lr=g["lr"], betas=tuple(g["betas"]), eps=g["eps"],
                                    weight_decay=g["weight_decay"])


def define_optimizer(model, opt, process_group=None) -> FlatAdam:
    """optimizer.py:25-27 (Adam with default betas/eps: the --optim_* flags are not passed)."""
    return FlatAdam(model.parameters(), lr=opt.learning_rate, weight_decay=opt.weight_decay,
                    grad_clip=opt.grad_clip, process_group=process_group)


def define_joint_optimizers(model, opt, process_group=None):
    """The nesting of optimizer.py:49-95: one Adam per agent; outside REINFORCE both step in the
    speaker turn and the listener turn is dropped."""
    spk = define_optimizer(model.caption_generator, opt, process_group)
    lis = define_optimizer(model.vse, opt, process_group)
    if opt.retrieval_reward == "reinforce":
        return {"speaker": spk, "listener": lis}
    return {"speaker": {"speaker": spk, "listener": lis}}


def zeroing_optimizer(opt, optimizer_dict, optimizer):
    """optimizer.py:224-230."""
    if opt.retrieval_reward != "reinforce" and opt.is_alternating:
        for agent in optimizer_dict["speaker"].keys():
            optimizer_dict["speaker"][agent].zero_grad()
    else:
        optimizer.zero_grad()


def update_optimizer(optimizer_dict, optimizer, opt):
    """optimizer.py:233-242: clip_gradient (elementwise clamp) then step, for both agents outside
    REINFORCE."""
    if opt.retrieval_reward != "reinforce" and opt.is_alternating:
        for agent in optimizer_dict["speaker"].keys():
            optimizer_dict["speaker"][agent].step(grad_clip=opt.grad_clip)
    else:
        optimizer.step(grad_clip=opt.grad_clip)
