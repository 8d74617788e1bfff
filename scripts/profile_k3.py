"""Driver for ncu: a few bf16 PPO minibatch-gradient calls at the benchmark shape (mb=32768)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from stoix_b200 import ops
B, mb, D, A = 524288, 32768, 64, 8
dev = "cuda:0"
sa, sc = ops.MlpSpec((D, 256, 256, A)), ops.MlpSpec((D, 256, 256, 1))
_, coff, total = ops.arena_offsets(sa, sc)
g = torch.Generator(device=dev).manual_seed(0)
arena = torch.randn(total, device=dev, generator=g) * 0.05
shadow = ops.cast_bf16(arena)
obs = torch.randn(B, D, device=dev, generator=g).to(torch.bfloat16)
batch = ops.PpoBatch(obs, torch.randint(0, A, (B,), device=dev, dtype=torch.int32, generator=g), -torch.rand(B, device=dev, generator=g) - 1.0,
                     torch.randn(B, device=dev, generator=g), torch.randn(B, device=dev, generator=g), torch.randn(B, device=dev, generator=g),
                     adv_stats=torch.tensor([0.0, 1.0], device=dev), perm=ops.make_permutation(B, 1, 0, device=dev))
grads, metrics = torch.zeros(total, device=dev), torch.zeros(8, device=dev)
ws = ops.ppo_workspace(sa, sc, mb, ops.STX_PREC_BF16, dev)
clk = torch.zeros(64, dtype=torch.int64, device=dev)
from stoix_b200 import _lib
_lib.load().stx_tc_debug_set_clock_buffer(clk.data_ptr())
for i in range(4):
    ops.ppo_minibatch_grads(sa, sc, arena, batch, i * mb, mb, 0.2, 0.01, 0.5, True, grads, metrics, ws, precision=ops.STX_PREC_BF16, param_arena_bf16=shadow)
torch.cuda.synchronize()
c = clk.cpu().tolist()
base = min(x for x in c if x > 0)
names = {0: "mma: tile start", 1: "mma: x_full+prev E4 ok -> issue G0", 2: "mma: G0 issued, wait E0", 3: "mma: E0 done -> issue G1", 4: "mma: wait E1",
         5: "mma: E1 done -> issue G2", 6: "mma: wait E2", 7: "mma: E2 done -> issue G3", 8: "mma: wait E3", 9: "mma: E3 done -> issue G4",
         16: "epi: wait G0", 17: "epi: G0 done", 18: "epi: wait G1 (E0 finished)", 19: "epi: G1 done", 20: "epi: wait G2 (E1 finished)", 21: "epi: G2 done",
         22: "epi: wait G3 (E2 finished)", 40: "E2: logits loaded", 41: "E2: loss math done", 42: "E2: butterfly done", 43: "E2: stores issued", 44: "E2: fence.proxy.async done", 48: "prod: loads issued (tile 1)", 49: "prod: x_empty ok", 50: "prod: smem+xg stores issued", 51: "prod: fence done", 23: "epi: G3 done", 24: "epi: wait G4 (E3 finished)", 25: "epi: G4 done", 26: "epi: E3 arrive", 27: "epi: E4 arrive"}
for k in sorted(names, key=lambda k: c[k]):
    if c[k] > 0:
        print(f"{c[k] - base:8d} cyc  {names[k]}")
