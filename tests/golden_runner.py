"""Replays tests/golden/multistep_vectors.json against any implementation of
batch_truncated_generalized_advantage_estimation (oracle on CPU, CUDA kernel on GPU)."""
import json
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).parent / "golden" / "multistep_vectors.json"


def load_cases():
    return json.loads(GOLDEN.read_text())["cases"]


def run_case(case, gae_fn):
    """gae_fn(r_t, discount_t, lambda_, values=None, v_tm1=None, v_t=None, truncation_t=None,
    time_major=False) -> (advantages, targets) as numpy arrays (batch-major unless time_major)."""
    inp = {k: (np.asarray(v, np.float32) if isinstance(v, list) else v) for k, v in case["inputs"].items()}
    lam = inp.pop("lambda_")
    r_t, discount_t = inp.pop("r_t"), inp.pop("discount_t")
    adv, tgt = gae_fn(r_t, discount_t, lam, **inp)
    adv, tgt = np.asarray(adv), np.asarray(tgt)
    assert adv.shape == r_t.shape and tgt.shape == r_t.shape
    v_tm1 = inp["values"][:, :-1] if "values" in inp else inp["v_tm1"]
    for chk in case["checks"]:
        kind, atol = chk["kind"], chk.get("atol", 1e-6)
        if kind == "adv_allclose":
            np.testing.assert_allclose(adv, np.asarray(chk["expected"], np.float32), atol=atol)
        elif kind == "targets_are_vtm1_plus":
            np.testing.assert_allclose(tgt, np.asarray(chk["expected"], np.float32) + v_tm1, atol=atol)
        elif kind == "targets_are_vtm1_plus_adv":
            np.testing.assert_allclose(tgt, v_tm1 + adv, atol=atol)
        elif kind == "adv_at":
            i, j = chk["index"]
            np.testing.assert_allclose(adv[i, j], chk["expected"], atol=atol)
        elif kind == "same_as_split_interface":
            vals = inp["values"]
            kw = {k: v for k, v in inp.items() if k != "values"}
            a2, t2 = gae_fn(r_t, discount_t, lam, v_tm1=np.ascontiguousarray(vals[:, :-1]), v_t=np.ascontiguousarray(vals[:, 1:]), **kw)
            np.testing.assert_allclose(adv, a2, atol=atol)
            np.testing.assert_allclose(tgt, t2, atol=atol)
        elif kind == "same_with_array_lambda":
            a2, t2 = gae_fn(r_t, discount_t, np.full_like(discount_t, lam), **inp)
            np.testing.assert_allclose(adv, a2, atol=atol)
            np.testing.assert_allclose(tgt, t2, atol=atol)
        elif kind == "same_as_values":
            a2, _ = gae_fn(r_t, discount_t, lam, values=np.asarray(chk["values"], np.float32))
            np.testing.assert_allclose(adv, a2, atol=atol)
        elif kind == "same_time_major":
            kw = {k: np.ascontiguousarray(v.T) for k, v in inp.items()}
            a2, t2 = gae_fn(np.ascontiguousarray(r_t.T), np.ascontiguousarray(discount_t.T), lam, time_major=True, **kw)
            np.testing.assert_allclose(adv, np.asarray(a2).T, atol=atol)
            np.testing.assert_allclose(tgt, np.asarray(t2).T, atol=atol)
        else:
            raise AssertionError(f"unknown check {kind}")
