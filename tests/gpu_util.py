"""Helpers shared by the GPU parity tests (CUDA path vs. the CPU oracle on the same seeded inputs)."""
import torch

from oracle import synth


def cuda_params(P):
    return {k: v.detach().clone().cuda().contiguous() for k, v in P.items()}


def pack_keep(keep_padded, att_masks):
    """[B, L, R] float keep-mask (oracle layout) -> uint8 [NL, R] over the valid regions."""
    if att_masks is None:
        return keep_padded.reshape(-1, keep_padded.shape[-1]).to(torch.uint8).cuda().contiguous()
    return keep_padded[att_masks > 0].to(torch.uint8).cuda().contiguous()


def u8(t):
    return None if t is None else t.to(torch.uint8).cuda().contiguous()


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30)


REAL = synth.Dims()
