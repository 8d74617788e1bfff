"""Pins oracle/ppo_oracle.gae against every golden vector of the reference's own unit test
(stoix/tests/multistep_test.py, transcribed in tests/golden/)."""
import numpy as np
import pytest

from oracle import ppo_oracle as O
from tests.golden_runner import load_cases, run_case


@pytest.mark.parametrize("case", load_cases(), ids=lambda c: c["name"])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_oracle_matches_reference_vectors(case, dtype):
    def fn(r_t, discount_t, lambda_, **kw):
        return O.gae(r_t, discount_t, lambda_, dtype=dtype, **kw)

    run_case(case, fn)


def test_fp64_oracle_vs_hand_tables_residual():
    """the hand tables are rounded; the exact recurrence differs by < 1e-3 from them."""
    case = [c for c in load_cases() if c["name"] == "basic_gae_lambda_0.4"][0]
    inp = case["inputs"]
    adv, _ = O.gae(np.array(inp["r_t"]), np.array(inp["discount_t"]), 0.4, values=np.array(inp["values"]))
    exp = np.array(case["checks"][0]["expected"])
    assert np.abs(adv - exp).max() < 1e-3


def test_standardize_matches_definition():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((16, 8)) * 3 + 1
    y = O.standardize(x)
    assert abs(y.mean()) < 1e-12
    np.testing.assert_allclose(y, (x - x.mean()) / np.sqrt(x.var() + 1e-5), rtol=1e-12)
