"""Driver for `ncu --set full`: one forward + backward of the GRU and LSTM sequence kernels at the rec_ppo minibatch shape
(T = 256 steps, 128 sequences, H = 128), after a warm-up of each.
    ncu --set full --clock-control none --nvtx --nvtx-include "measure/" -o /tmp/rep/rec python scripts/profile_rec_kernels.py"""
import sys

import torch

sys.path.insert(0, ".")
from stoix_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
T, E, H = 256, 128, 128
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g)
reset = (torch.rand(T, E, device=dev, generator=g) < 0.01).to(torch.uint8)
gi3, gi4 = rn(T, E, 3 * H), rn(T, E, 4 * H)
wh3, wh4, bhn = rn(H, 3 * H) * 0.1, rn(H, 4 * H) * 0.1, rn(H) * 0.1
h0, c0 = rn(E, H), rn(E, 2 * H)
d_h = rn(T, E, H)
ws_g, ws_l = ops.gru_workspace(T, E, H, dev), ops.lstm_workspace(T, E, H, dev)
d_gi3, d_gi4 = torch.zeros(T, E, 3 * H, device=dev), torch.zeros(T, E, 4 * H, device=dev)
dw3, dw4, db = torch.zeros(H, 3 * H, device=dev), torch.zeros(H, 4 * H, device=dev), torch.zeros(H, device=dev)


def all_ops():
    ops.gru_sequence_forward(gi3, reset, h0, wh3, bhn, ws_g)
    ops.gru_sequence_backward(d_h, reset, wh3, ws_g, d_gi3, d_w_h=dw3, d_b_hn=db)
    ops.lstm_sequence_forward(gi4, reset, c0, wh4, ws_l)
    ops.lstm_sequence_backward(d_h, reset, wh4, ws_l, d_gi4, d_w_h=dw4)


all_ops()
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("measure")
all_ops()
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
