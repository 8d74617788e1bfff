"""bf16 tensor-core path (tcgen05): fused MLP forward vs the oracle with bf16-rounded GEMM operands
(weights, inputs and hidden activations rounded to bf16, fp32 accumulate -- oracle.mlp_forward(
bf16_operands=True)).  Tolerance: the only differences left are fp32 accumulation order and rare
1-ulp bf16 rounding flips of a hidden activation: rtol 2e-3 / atol 2e-3 on outputs; vs the pure fp32
reference the bf16 path is reported at rtol 2e-2."""
import numpy as np
import pytest
import torch

from oracle import ppo_oracle as O

pytestmark = pytest.mark.gpu


def _t(x, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype, device="cuda:0")


def _net(rng, D, A, head_scale):
    p = O.init_mlp(rng, [D, 256, 256, A], head_scale)
    for i in range(3):
        p.b[i] = rng.standard_normal(p.b[i].shape) * 0.1
    p.W[-1] = rng.standard_normal(p.W[-1].shape) * 0.3
    return p


@pytest.mark.parametrize("M,D,A", [(128, 64, 8), (4096, 64, 8), (300, 64, 1), (128 * 150 + 5, 64, 8), (256, 32, 16), (64, 16, 2)])
def test_tc_forward_matches_bf16_oracle(M, D, A):
    from stoix_b200 import ops

    rng = np.random.default_rng(M + A)
    net = _net(rng, D, A, 1.0)
    x = rng.standard_normal((M, D)).astype(np.float32)
    spec = ops.MlpSpec((D, 256, 256, A))
    params = _t(net.flat())
    shadow = ops.cast_bf16(params)
    xb = _t(x).to(torch.bfloat16)
    out, h1, h2 = ops.tc_debug_forward(spec, params, shadow, xb)
    torch.cuda.synchronize()
    ref, acts = O.mlp_forward(net.astype(np.float32), x, bf16_operands=True)
    e1 = np.abs(h1.cpu().numpy() - acts[1]).max()
    e2 = np.abs(h2.cpu().numpy() - acts[2]).max()
    e3 = np.abs(out.cpu().numpy() - ref).max()
    print(f"M={M} D={D} A={A}: max|h1 err|={e1:.3e} max|h2 err|={e2:.3e} max|out err|={e3:.3e}")
    np.testing.assert_allclose(h1.cpu().numpy(), acts[1], rtol=1e-2, atol=1e-2)   # a bf16 ulp at |h|~2 is 1.6e-2 * 0.5
    np.testing.assert_allclose(h2.cpu().numpy(), acts[2], rtol=1e-2, atol=2e-2)
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=2e-3, atol=5e-3)
    # the public entry point (no debug outputs) gives the same numbers
    out2 = ops.mlp_forward(spec, params, xb, precision=ops.STX_PREC_BF16, params_bf16=shadow)
    assert torch.equal(out, out2)
    # and stays within the stated bf16-vs-fp32 band of the pure fp32 reference
    ref32, _ = O.mlp_forward(net, x.astype(np.float64))
    assert np.abs(out.cpu().numpy() - ref32).max() < 2e-2 * max(1.0, np.abs(ref32).max())
