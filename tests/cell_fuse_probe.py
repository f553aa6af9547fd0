"""Helper of test_gpu_cell_fuse.py: one free-running ST-Gumbel decode + listener forward at the real
model size with Philox noise and dropout, every saved tensor of both passes dumped to the path in
argv[1].  Run twice (with and without COOPCAP_CELL_FUSE=1) the dumps must be bit-identical."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import synth  # noqa: E402  (parameter / batch generators only)
from gpu_util import REAL, cuda_params  # noqa: E402
from cooperativeimagecaptioning_b200 import engine as EN  # noqa: E402

B, L, seed = int(sys.argv[2]), 9, 77
d = REAL
T, V = d.seq_length, d.vocab_size
Ps = synth.speaker_params(d, seed=seed, eos_bias=-3.0)
Pl = synth.listener_params(d, seed=seed + 1)
batch = synth.make_batch(d, B, L, seed + 2, varlen=True, min_regions=2)
Pc, Plc = cuda_params(Ps), cuda_params(Pl)
packed_s, packed_l = EN.PackedSpeaker().get(Pc), EN.PackedListener().get(Plc)
off, NL = EN.region_offsets(batch.att_masks.cuda(), B, L)
rnd = EN.SpeakerRandom(seed=12345, drop_p=0.5)
sp = EN.speaker_forward(Pc, packed_s, batch.att_feats.cuda(), off, NL, n_steps=T, mode=EN.MODE_ST_GUMBEL,
                        inv_tau=1.0, start_token=V + 1, rnd=rnd)
tok_sb = torch.cat([torch.full((1, B), V + 1, dtype=torch.int64, device="cuda"), sp.t["tok_out"]], 0).contiguous()
lp = EN.listener_forward(Plc, packed_l, batch.fc_feats.cuda(), tok_sb, sp.t["cap_len"])
torch.cuda.synchronize()
out = {}
for name in ("tok_out", "logp", "c_all", "u_all", "out16", "cap_len", "s_all"):
    out["sp." + name] = sp.t[name].cpu()
E = d.input_encoding_size
out["sp.xh16_h"] = sp.t["xh16"][:, :, E:].cpu()          # h_0 .. h_T
out["sp.xh16_x"] = sp.t["xh16"][:T, :, :E].cpu()         # x_0 .. x_{T-1} (slot T's x is never written)
for name in ("h32", "h16", "gates", "scores", "loss"):
    out["lp." + name] = lp.t[name].cpu()
torch.save(out, sys.argv[1])
