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
for i in range(4):
    ops.ppo_minibatch_grads(sa, sc, arena, batch, i * mb, mb, 0.2, 0.01, 0.5, True, grads, metrics, ws, precision=ops.STX_PREC_BF16, param_arena_bf16=shadow)
torch.cuda.synchronize()
