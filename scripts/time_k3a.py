"""Times the K3 launches alone (CUDA events)."""
import sys
import torch
sys.path.insert(0, ".")
exec(open("scripts/profile_k3.py").read().split("clk = torch.zeros")[0])
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for _ in range(3):
    ops.ppo_minibatch_grads(sa, sc, arena, batch, 0, mb, 0.2, 0.01, 0.5, True, grads, metrics, ws, precision=ops.STX_PREC_BF16, param_arena_bf16=shadow)
torch.cuda.synchronize()
ev[0].record()
for i in range(16):
    ops.ppo_minibatch_grads(sa, sc, arena, batch, i * mb, mb, 0.2, 0.01, 0.5, True, grads, metrics, ws, precision=ops.STX_PREC_BF16, param_arena_bf16=shadow)
ev[1].record()
torch.cuda.synchronize()
print(f"K3 (fwd/bwd + dW + reduce) per minibatch: {ev[0].elapsed_time(ev[1]) / 16 * 1000:.1f} us")
