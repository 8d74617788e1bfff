"""Per-CTA cycle accounting of the fused K3 launch at the benchmark shape (mb = 32768): who waits for what.

    python scripts/profile_k3_fused.py            (honours STX_K3_FUSED / STX_K3_SPLIT)

Prints, per role, min / median / max over its CTAs of the counters the kernel leaves in the profile buffer
(stx_tc_debug_set_prof_buffer), the CTA-0 timeline stamps of the K3a role, and CUDA-event timings of the call."""
import statistics
import sys

import torch

sys.path.insert(0, ".")
from stoix_b200 import _lib, ops  # noqa: E402

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
prof = torch.zeros(148 * 8, dtype=torch.int64, device=dev)
lib = _lib.load()


def call(i):
    ops.ppo_minibatch_grads(sa, sc, arena, batch, (i % 16) * mb, mb, 0.2, 0.01, 0.5, True, grads, metrics, ws, precision=ops.STX_PREC_BF16,
                            param_arena_bf16=shadow, overwrite=True)


for i in range(3):
    call(i)
torch.cuda.synchronize()
ts = []
for i in range(8):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    call(i)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e3)
print(f"K3 call (all launches of one minibatch step), us: median {statistics.median(ts):.1f}  min {min(ts):.1f}  max {max(ts):.1f}")
lib.stx_tc_debug_set_prof_buffer(prof.data_ptr())
call(3)
torch.cuda.synchronize()
lib.stx_tc_debug_set_prof_buffer(None)
p = prof.view(148, 8).cpu().tolist()
roles = {}
for row in p:
    roles.setdefault(row[0], []).append(row)
names = {1: ("K3a CTA", ["total", "in flag_signal (warp 5)", "tiles"]),
         }
for role in sorted(roles):
    rows = roles[role]
    if role == 0:
        print(f"role 0 (no record): {len(rows)} CTAs")
        continue
    cols = ["total", "flag/signal", "stage-free wait", "full-stage wait (MMA)", "iters"] if role >= 2 else ["total", "in flag_signal (warp 5)", "tiles"]
    label = "K3a CTA" if role == 1 else f"dW job {role - 2}"
    print(f"{label}: {len(rows)} CTAs")
    for ci, cname in enumerate(cols):
        v = [r[1 + ci] for r in rows]
        print(f"    {cname:28s} min {min(v):9d}  median {int(statistics.median(v)):9d}  max {max(v):9d}")
