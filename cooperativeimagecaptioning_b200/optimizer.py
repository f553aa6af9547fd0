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

import os
from typing import Dict, Iterable, List, Optional

import torch

from . import engine as EN


# Early exchange of gradient slices (FlatAdam.reduce_async) is OPT-IN (COOPCAP_EARLY_REDUCE=1).
# Measured with tools/dp_probe.py (r2): on 2 GPUs the listener + logit slices exchanged under the
# speaker's BPTT give 5.830 ms/step against 5.843 for one all-reduce after backward (5.565 without any
# exchange) -- the NCCL kernels take from the BPTT kernels what they hide; on 8 GPUs (NVLS, 24
# channels) four slices cost 5.989 ms/step against 5.891 for the single 104.5 MB all-reduce.
_EARLY_REDUCE = os.environ.get("COOPCAP_EARLY_REDUCE", "0") == "1"


class FlatAdam:
    """Adam over flat fp32 buckets (see the module docstring).

    A parameter that already lives in ANOTHER FlatAdam's bucket -- the embedding shared by speaker
    and listener under --share_embed (AlternatingJointModel.py:85-88): it is in both agents'
    `parameters()`, so the reference builds two torch Adams that both update it with their own
    moments (optimizer.py:233-242) -- is *aliased*, not re-bucketed: it stays a view of its first
    owner's parameter bucket, and this optimizer keeps its own gradient / moment segments for it at
    the tail of its buckets ("foreign" segments) and updates it through the owner's storage."""

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
        self.foreign = [getattr(p, "_coopcap_bucket", None) is not None for p in self.params]
        self.offsets, n = [0] * len(self.params), 0
        for own_pass in (True, False):               # own segments first, foreign ones at the tail
            for i, p in enumerate(self.params):
                if self.foreign[i] != own_pass:
                    self.offsets[i] = n
                    n += (p.numel() + 3) // 4 * 4
            if own_pass:
                self.own_numel = n
        self.numel = n
        self.flat_param = torch.zeros(self.own_numel, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self._grad_views = []
        for p, o, far in zip(self.params, self.offsets, self.foreign):
            if not far:
                view = self.flat_param[o:o + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view                                # parameters become views of the bucket
                p._coopcap_bucket = self
            elif not (p.data.is_contiguous() and p.data.dtype == torch.float32):
                raise ValueError("FlatAdam: an aliased parameter must be contiguous fp32")
            self._grad_views.append(self.flat_grad[o:o + p.numel()].view_as(p))
            p.grad = None
        self.param_groups = [dict(lr=lr, weight_decay=weight_decay, betas=betas, eps=eps,
                                  params=self.params)]
        self.grad_clip = grad_clip
        self.step_count = 0
        self.process_group = process_group
        self._reduced = False          # this step's gradient bucket has been summed over the ranks
        self._tail_reduced = False     # ... and the foreign tail was copied from a summed bucket
        self._index = {id(p): i for i, p in enumerate(self.params)}
        self._early = []               # bucket ranges already all-reduced by reduce_async this step
        self._comm = None              # side stream of the early exchanges

    # reference call shape: optimizer.zero_grad()
    def zero_grad(self, set_to_none: bool = True):
        """Gradients are dropped, not zero-filled: the first gradient a parameter receives in the
        next backward pass is then adopted by autograd as `p.grad` without an accumulation kernel
        (24 read-modify-write passes over the 104 MB bucket per step otherwise); `gather_grads`
        moves them into the flat bucket with one multi-tensor copy."""
        for p in self.params:
            p.grad = None
        self._reduced = self._tail_reduced = False
        self._early = []

    def gather_grads(self):
        """p.grad of every parameter -> its segment of the flat gradient bucket (zeros where a
        parameter received no gradient); afterwards p.grad IS that segment."""
        src, dst = [], []
        tail_from_reduced = []
        for p, v, far in zip(self.params, self._grad_views, self.foreign):
            g = p.grad
            if g is None:
                v.zero_()
            elif g.data_ptr() != v.data_ptr():
                src.append(g.detach().to(torch.float32).reshape(v.shape))
                dst.append(v)
            if far:
                # the other owner of a shared parameter may already have summed this gradient
                # over the ranks (it then sits in that optimizer's bucket): do not sum it twice
                other = getattr(g, "_coopcap_owner", None) if g is not None else None
                tail_from_reduced.append(other is not None and other is not self and other._reduced)
            p.grad = v
        if src:
            torch._foreach_copy_(dst, src)
        if tail_from_reduced:
            if any(tail_from_reduced) != all(tail_from_reduced):
                raise RuntimeError("FlatAdam: shared parameters were reduced inconsistently")
            self._tail_reduced = all(tail_from_reduced)
        for v in self._grad_views:
            v._coopcap_owner = self

    # ---- gradients written straight into the bucket, slices reduced while backward still runs ----
    def grad_view(self, p) -> Optional[torch.Tensor]:
        """This optimizer's gradient-bucket view for parameter `p` (None if `p` is not ours).  The
        fused backward nodes write a parameter's gradient there directly and set `p.grad` to it, so
        neither autograd's accumulation nor gather_grads' copy touches the 104 MB again."""
        i = self._index.get(id(p))
        return None if i is None else self._grad_views[i]

    def _world(self):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(self.process_group)
        return 1

    def reduce_async(self, params) -> bool:
        """All-reduce (sum) the bucket range that holds the gradients of `params` NOW, on a side
        stream, while the caller keeps enqueueing backward work: the listener's 46.7 MB are final
        before the speaker's ~2.5 ms BPTT starts, so their exchange is hidden behind it.  The
        gradients must already sit in the bucket (grad_view) and be final.  Returns False (and does
        nothing) when the parameters do not form one contiguous bucket range of their own."""
        if self._world() <= 1 or self.flat_grad.device.type != "cuda" or not _EARLY_REDUCE:
            return False
        import torch.distributed as dist
        idx = sorted(self._index[id(p)] for p in params if id(p) in self._index)
        if not idx or any(self.foreign[i] for i in idx):
            return False
        lo = min(self.offsets[i] for i in idx)
        hi = max(self.offsets[i] + (self.params[i].numel() + 3) // 4 * 4 for i in idx)
        inside = {i for i in range(len(self.params))
                  if not self.foreign[i] and lo <= self.offsets[i] < hi}
        if inside != set(idx) or any(a < hi and lo < b for a, b in self._early):
            return False
        if self._comm is None:
            self._comm = torch.cuda.Stream(device=self.flat_grad.device)
        self._comm.wait_stream(torch.cuda.current_stream(self.flat_grad.device))
        with torch.cuda.stream(self._comm):
            dist.all_reduce(self.flat_grad[lo:hi], op=dist.ReduceOp.SUM, group=self.process_group)
        self._early.append((lo, hi))
        return True

    def all_reduce(self):
        """Sum the gradient bucket over the data-parallel ranks (no-op for a single process);
        ranges already exchanged by reduce_async are skipped."""
        import torch.distributed as dist
        self.gather_grads()
        n = self._world()
        if n > 1 and not self._reduced:
            end = self.own_numel if self._tail_reduced else self.numel
            todo, at = [], 0
            for lo, hi in sorted(self._early):
                if lo > at:
                    todo.append((at, lo))
                at = max(at, hi)
            if at < end:
                todo.append((at, end))
            for lo, hi in todo:
                dist.all_reduce(self.flat_grad[lo:hi], op=dist.ReduceOp.SUM, group=self.process_group)
            if self._early and self._comm is not None:       # the early slices must have landed
                torch.cuda.current_stream(self.flat_grad.device).wait_stream(self._comm)
        self._reduced = True
        return n

    def step(self, grad_clip: Optional[float] = None, world_size: Optional[int] = None):
        """clamp + Adam; `world_size` > 1 divides the (already all-reduced) gradient first."""
        if self.flat_grad.device.type != "cuda":
            raise EN._lib.CoopcapError("FlatAdam.step needs CUDA parameters (there is no CPU path)")
        if world_size is None:
            world_size = self.all_reduce()
        else:
            self.gather_grads()
        g = self.param_groups[0]
        self.step_count += 1
        clip = self.grad_clip if grad_clip is None else grad_clip
        kw = dict(step=self.step_count, lr=g["lr"], grad_scale=1.0 / world_size, clip=clip,
                  beta1=g["betas"][0], beta2=g["betas"][1], eps=g["eps"],
                  weight_decay=g["weight_decay"])
        if self.own_numel:
            k = self.own_numel
            EN.clamp_adam_(self.flat_param, self.flat_grad[:k], self.exp_avg[:k], self.exp_avg_sq[:k], **kw)
        for p, o, far in zip(self.params, self.offsets, self.foreign):
            if far:       # shared parameter: our moments and gradient copy, the first owner's storage
                k = p.numel()
                EN.clamp_adam_(p.data.view(-1), self.flat_grad[o:o + k], self.exp_avg[o:o + k],
                               self.exp_avg_sq[o:o + k], **kw)
        # the kernel wrote the parameters through raw pointers: tell the packed bf16 operand
        # caches (engine.PackedSpeaker / PackedListener) that the masters changed
        EN.bump_weights_epoch()

    # torch.optim-compatible checkpoint layout (optimizer.py:191-221 saves optimizer.state_dict())
    def state_dict(self) -> Dict:
        state = {}
        if self.step_count > 0:     # torch.optim.Adam has no per-parameter state before its first step
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
        """Accepts what `torch.optim.Adam.state_dict()` (the reference's checkpoints) or
        `FlatAdam.state_dict()` produced: one parameter group, state keyed by parameter position."""
        groups = sd["param_groups"]
        if len(groups) != 1 or len(groups[0]["params"]) != len(self.params):
            raise ValueError("loaded state dict has a different number of parameter groups / "
                             "parameters than this optimizer")
        if groups[0].get("amsgrad", False):
            raise ValueError("amsgrad checkpoints are not supported (the reference never sets it)")
        index = {pid: i for i, pid in enumerate(groups[0]["params"])}
        for pid, st in sd["state"].items():
            i = index[pid]
            p, o = self.params[i], self.offsets[i]
            n = p.numel()
            if st["exp_avg"].numel() != n:
                raise ValueError(f"optimizer state of parameter {i} has {st['exp_avg'].numel()} "
                                 f"elements, the parameter has {n}")
            self.exp_avg[o:o + n].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
            self.step_count = max(int(float(st["step"])), 0)
        g = groups[0]
        self.param_groups[0].update(lr=g["lr"], betas=tuple(g["betas"]), eps=g["eps"],
                                    weight_decay=g["weight_decay"])


def define_optimizer(model, opt, process_group=None) -> FlatAdam:
    """optimizer.py:25-27 (Adam with default betas/eps: the --optim_* flags are not passed)."""
    return FlatAdam(model.parameters(), lr=opt.learning_rate, weight_decay=opt.weight_decay,
                    grad_clip=opt.grad_clip, process_group=process_group)


# ------------------------------------------------------------------ checkpoint resume / save
def load_optimizer_path(opt, curr_turn=None):
    """optimizer.py:9-22: `<start_from>/<turn>_optimizer.pth` (joint training, None when the file
    is missing) or `<start_from>/optimizer.pth` (single-agent phases)."""
    if opt.is_alternating:
        path = os.path.join(opt.start_from, curr_turn + "_optimizer.pth")
        return path if os.path.isfile(path) else None
    if opt.start_from is None:
        return None
    return os.path.join(opt.start_from, "optimizer.pth")


def load_state_dict(optimizer, optimizer_path, agent=""):
    """optimizer.py:30-40."""
    sd = torch.load(optimizer_path, map_location=None if torch.cuda.is_available() else "cpu")
    optimizer.load_state_dict(sd)
    print(f"\n Loaded {agent} optimizer from {optimizer_path} \n")
    return optimizer


def load_optimizer_from_checkpoint(opt, curr_turn, optimizer):
    """optimizer.py:43-46."""
    return load_state_dict(optimizer, load_optimizer_path(opt, curr_turn), curr_turn)


def _resume_joint(optimizer, opt, turn, resume, fallback_path):
    """Joint-training resume order (optimizer.py:53-64, 74-86): this turn's own checkpoint if it
    exists, else -- unless the embedding is shared -- the optimizer of the phase the agent was
    pre-trained in; a fresh optimizer otherwise."""
    if not resume:
        return optimizer, False
    if load_optimizer_path(opt, turn):
        return load_optimizer_from_checkpoint(opt, turn, optimizer), True
    if not opt.share_embed:
        return load_state_dict(optimizer, fallback_path(), turn), True
    return optimizer, False


def define_speaker_optimizer_joint_training(model, opt, start_from_exist, optimizer_dict, curr_turn):
    """optimizer.py:49-67."""
    optimizer, loaded = _resume_joint(define_optimizer(model.caption_generator, opt), opt, curr_turn,
                                      start_from_exist, lambda: opt.speaker_stage_2_optimizer_path)
    if not start_from_exist:
        print('Loaded new "speaker" optimizer')
    optimizer_dict[curr_turn] = optimizer
    return optimizer_dict


def define_listener_optimizer_joint_training(model, opt, start_from_exist, optimizer_dict, curr_turn):
    """optimizer.py:70-96.  Outside REINFORCE the listener's optimizer is nested under the speaker
    turn (both agents step on the same loss) and the listener turn is removed from
    opt.alternating_turn.  One deliberate difference: the reference registers the listener's
    optimizer only when `start_from` is set and then fails in train.py (KeyError at :493 /
    AttributeError in zeroing_optimizer) on a run without a snapshot; here a fresh run gets a fresh
    listener optimizer."""
    optimizer, loaded = _resume_joint(
        define_optimizer(model.vse, opt), opt, curr_turn, start_from_exist,
        lambda: os.path.join(os.path.split(opt.initialize_retrieval)[0], "optimizer.pth"))
    if not loaded:
        print('\n Using new "listener" optimizer \n')
    if opt.retrieval_reward == "reinforce":
        optimizer_dict[curr_turn] = optimizer
    else:
        optimizer_dict["speaker"] = {"speaker": optimizer_dict["speaker"], "listener": optimizer}
        opt.alternating_turn.remove("listener")
    return optimizer_dict


def _single_agent(agent_module, opt, resume_path, optimizer_dict, agent=""):
    optimizer = define_optimizer(agent_module, opt)
    if resume_path is not None:
        optimizer = load_state_dict(optimizer, resume_path, agent)
    optimizer_dict["optimizer"] = optimizer
    return optimizer_dict


def define_pretraining_listener_optimizer(model, opt, start_from_exist, optimizer_dict, optimizer_exist):
    """optimizer.py:99-111 (phase 1: the listener alone on ground-truth captions)."""
    path = os.path.join(opt.start_from, "optimizer.pth") if (start_from_exist and optimizer_exist) else None
    return _single_agent(model.vse, opt, path, optimizer_dict)


def define_pretraining_speaker_optimizer(model, opt, start_from_exist, optimizer_dict, optimizer_exist):
    """optimizer.py:114-126 (phase 2: speaker MLE)."""
    path = os.path.join(opt.start_from, "optimizer.pth") if (start_from_exist and optimizer_exist) else None
    return _single_agent(model.caption_generator, opt, path, optimizer_dict)


def define_only_speaker_optimizer(model, opt, start_from_exist, optimizer_dict, optimizer_exist):
    """optimizer.py:129-146 (phase 3: speaker fine-tuning against a frozen listener)."""
    path = None
    if start_from_exist:
        if optimizer_exist:
            path = os.path.join(opt.start_from, "optimizer.pth")
        elif not opt.share_embed:
            path = opt.speaker_stage_2_optimizer_path
    return _single_agent(model.caption_generator, opt, path, optimizer_dict, "speaker")


def load_optimizer(model, opt):
    """optimizer.py:150-188: the optimizer dictionary train.py drives -- one entry per alternating
    turn (`{'speaker': ..., 'listener': ...}` for REINFORCE, `{'speaker': {'speaker': ...,
    'listener': ...}}` otherwise when resuming), or `{'optimizer': ...}` for phases 1-3."""
    start_from_exist = vars(opt).get("start_from", None) is not None
    optimizer_dict = {}
    if opt.is_alternating:
        joint = {"speaker": define_speaker_optimizer_joint_training,
                 "listener": define_listener_optimizer_joint_training}
        for curr_turn in list(opt.alternating_turn):      # the listener builder may edit the list
            if curr_turn in joint:
                optimizer_dict = joint[curr_turn](model, opt, start_from_exist, optimizer_dict, curr_turn)
        return optimizer_dict
    by_phase = {1: define_pretraining_listener_optimizer, 2: define_pretraining_speaker_optimizer,
                3: define_only_speaker_optimizer}
    if opt.phase in by_phase:
        optimizer_dict = by_phase[opt.phase](model, opt, start_from_exist, optimizer_dict,
                                             load_optimizer_path(opt))
    return optimizer_dict


def save_optimizer(opt, optimizer_dict):
    """optimizer.py:191-221: `<checkpoint_path>/<agent>_optimizer.pth` per agent in joint training
    (the nested gumbel / multinomial dictionary is flattened), `optimizer.pth` otherwise."""
    if not opt.is_alternating:
        todo = {"optimizer.pth": optimizer_dict["optimizer"]}
    elif opt.retrieval_reward == "reinforce":
        todo = {agent + "_optimizer.pth": o for agent, o in optimizer_dict.items()}
    else:
        todo = {agent + "_optimizer.pth": o for agent, o in optimizer_dict.get("speaker", {}).items()} \
            if isinstance(optimizer_dict.get("speaker"), dict) else {}
    for name, optimizer in todo.items():
        path = os.path.join(opt.checkpoint_path, name)
        torch.save(optimizer.state_dict(), path)
        print(f"\n Optimizer saved to {path} \n")


def clip_gradient(optimizer, grad_clip):
    """misc/utils.py:65-69 for foreign callers: with FlatAdam the clamp happens inside step()."""
    if isinstance(optimizer, FlatAdam):
        optimizer.grad_clip = grad_clip
        return
    for group in optimizer.param_groups:
        for p in group["params"]:
            if p.grad is not None:
                p.grad.data.clamp_(-grad_clip, grad_clip)


def set_lr(optimizer, lr):
    """misc/utils.py:60-62."""
    for group in optimizer.param_groups:
        group["lr"] = lr


def define_joint_optimizers(model, opt, process_group=None):
    """The nesting of optimizer.py:49-95 without a snapshot to resume from (bench / tests): one
    Adam per agent; outside REINFORCE both step in the speaker turn and the listener turn is
    dropped."""
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
