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
names = {0: "mma: tile start", 1: "mma: G0 issued (all parts)", 3: "mma: G1 issued (E0 finished)", 5: "mma: G2 issued (E1 finished)",
         7: "mma: G3 issued (E2 finished)", 9: "mma: G4 issued (E3 finished)",
         16: "epi: E0 start", 17: "epi: G0 part 0 ready", 18: "epi: E1 start", 19: "epi: G1 part 0 ready", 20: "epi: E2 start (E1 finished)",
         21: "epi: head ready", 41: "E2: loss math done", 44: "E2: dz in smem, head_done arrived", 22: "epi: E3 start", 24: "epi: E4 start",
         26: "epi: E3 finished", 27: "epi: E4 finished", 48: "prod: loads issued (tile 1)", 49: "prod: x_empty ok", 50: "prod: smem+xg stores issued",
         51: "prod: fence done"}
names.update({56: "kernel entry", 57: "setup done (barriers, TMEM, W2, biases)", 58: "mma: W0/W1 landed", 59: "mma: first X tile landed",
              61: "epi: all tiles done", 62: "kernel exit (thread 0)"})
for t in range(4):
    names.update({32 + 4 * t: f"tile {t}: E0 start", 33 + 4 * t: f"tile {t}: E2 start", 34 + 4 * t: f"tile {t}: E3 end", 35 + 4 * t: f"tile {t}: E4 end"})
for k in sorted(names, key=lambda k: c[k]):
    if c[k] > 0:
        print(f"{c[k] - base:8d} cyc  {names[k]}")
