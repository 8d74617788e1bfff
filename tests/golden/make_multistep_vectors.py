"""Writes tests/golden/multistep_vectors.json: the golden vectors of the reference's only unit test,
stoix/tests/multistep_test.py (GAE), transcribed case by case with the reference line numbers.

JAX is not installable in the build container, so the reference function itself
cannot be executed to produce outputs; what pins the oracle are the expected values the reference
authors wrote into their test (hand tables :83-96 at atol=1e-3, inline scalars elsewhere).  Each case
below is (inputs, list of checks); tests/test_oracle_golden.py replays them against oracle/ and
tests/test_gae_gpu.py against the CUDA kernel.

Run:  python tests/golden/make_multistep_vectors.py
"""
import json
from pathlib import Path

cases = []

# ---- shared data, multistep_test.py:36-74 -------------------------------------------------------
r_t = [[0.0, 0.0, 1.0, 0.0, -0.5], [0.0, 0.0, 0.0, 0.0, 1.0]]
values = [[1.0, 4.0, -3.0, -2.0, -1.0, -1.0], [-3.0, -2.0, -1.0, 0.0, 5.0, -1.0]]
discount_t = [[0.99, 0.99, 0.99, 0.99, 0.99], [0.9, 0.9, 0.9, 0.0, 0.9]]
expected = {  # multistep_test.py:83-96
    "1.0": [[-1.45118, -4.4557, 2.5396, 0.5249, -0.49], [3.0, 2.0, 1.0, 0.0, -4.9]],
    "0.7": [[-0.676979, -5.248167, 2.4846, 0.6704, -0.49], [2.2899, 1.73, 1.0, 0.0, -4.9]],
    "0.4": [[0.56731, -6.042, 2.3431, 0.815, -0.49], [1.725, 1.46, 1.0, 0.0, -4.9]],
}
for lam, exp in expected.items():
    cases.append({
        "name": f"basic_gae_lambda_{lam}", "ref": "multistep_test.py:113-150",
        "inputs": {"r_t": r_t, "discount_t": discount_t, "lambda_": float(lam), "values": values},
        "checks": [
            {"kind": "adv_allclose", "expected": exp, "atol": 1e-3},
            {"kind": "targets_are_vtm1_plus", "expected": exp, "atol": 1e-3},
            {"kind": "same_as_split_interface", "atol": 1e-6},
        ],
    })

# scalar vs array lambda, multistep_test.py:152-171
cases.append({
    "name": "scalar_vs_array_lambda", "ref": "multistep_test.py:152-171",
    "inputs": {"r_t": r_t, "discount_t": discount_t, "lambda_": 0.9, "values": values},
    "checks": [{"kind": "same_with_array_lambda", "atol": 1e-6}],
})

# truncation vs termination, multistep_test.py:175-223
cases.append({
    "name": "truncation_case", "ref": "multistep_test.py:186-209,219",
    "inputs": {"r_t": [[0.0, 0.0, 0.0, 0.0]], "discount_t": [[0.9, 0.9, 0.9, 0.9]], "lambda_": 1.0,
               "values": [[1.0, 1.0, 1.0, 1.0, 10.0]], "truncation_t": [[0.0, 0.0, 1.0, 0.0]]},
    "checks": [{"kind": "adv_at", "index": [0, 2], "expected": -0.1, "atol": 1e-5}],
})
cases.append({
    "name": "termination_case", "ref": "multistep_test.py:194-214,220",
    "inputs": {"r_t": [[0.0, 0.0, 0.0, 0.0]], "discount_t": [[0.9, 0.9, 0.0, 0.9]], "lambda_": 1.0,
               "values": [[1.0, 1.0, 1.0, 1.0, 10.0]]},
    "checks": [{"kind": "adv_at", "index": [0, 2], "expected": -1.0, "atol": 1e-5}],
})

# multiple truncations, multistep_test.py:225-258
cases.append({
    "name": "multiple_truncations", "ref": "multistep_test.py:225-258",
    "inputs": {"r_t": [[0.0, 1.0, 0.0, 0.0, 0.0, 1.0, 0.0]], "discount_t": [[0.9] * 7], "lambda_": 1.0,
               "values": [[0.0, 1.0, 2.0, 1.0, 0.0, 1.0, 0.0, 0.0]],
               "truncation_t": [[0.0, 0.0, 1.0, 0.0, 1.0, 0.0, 0.0]]},
    "checks": [
        {"kind": "adv_at", "index": [0, 6], "expected": 0.0, "atol": 5e-4},
        {"kind": "adv_at", "index": [0, 5], "expected": 0.0, "atol": 5e-4},
        {"kind": "adv_at", "index": [0, 4], "expected": 0.9, "atol": 5e-4},
        {"kind": "adv_at", "index": [0, 3], "expected": -0.19, "atol": 5e-3},
        {"kind": "adv_at", "index": [0, 2], "expected": -1.1, "atol": 5e-4},
    ],
})

# mixed truncation and termination, multistep_test.py:262-324
cases.append({
    "name": "mixed_truncation_and_termination", "ref": "multistep_test.py:262-324",
    "inputs": {"r_t": [[1.0, 0.0, 0.0], [0.0, 0.5, 0.0], [0.0, 0.0, 2.0]],
               "discount_t": [[0.9, 0.9, 0.0], [0.9, 0.9, 0.9], [0.9, 0.9, 0.0]], "lambda_": 1.0,
               "values": [[1.0, 1.0, 0.0, 0.0], [1.0, 1.0, 1.0, 0.0], [1.0, 1.0, 0.5, 0.0]],
               "truncation_t": [[0.0, 0.0, 0.0], [0.0, 0.0, 1.0], [0.0, 0.0, 0.0]]},
    "checks": [
        {"kind": "adv_at", "index": [0, 2], "expected": 0.0, "atol": 5e-4},
        {"kind": "adv_at", "index": [1, 2], "expected": -1.0, "atol": 5e-4},
        {"kind": "adv_at", "index": [2, 2], "expected": 1.5, "atol": 5e-4},
        {"kind": "targets_are_vtm1_plus_adv", "atol": 1e-3},
    ],
})

# values vs v_tm1/v_t interfaces, multistep_test.py:328-429
cases.append({
    "name": "autoreset_interface", "ref": "multistep_test.py:345-398,406-410",
    "inputs": {"r_t": [[1.0, 0.0, 0.0, 0.5, 0.0, 0.0]], "discount_t": [[0.9] * 6], "lambda_": 1.0,
               "v_tm1": [[1.0, 2.0, 3.0, 1.0, 1.5, 2.0]], "v_t": [[2.0, 3.0, 4.0, 1.5, 2.0, 1.0]],
               "truncation_t": [[0.0, 0.0, 1.0, 0.0, 0.0, 1.0]]},
    "checks": [
        {"kind": "adv_at", "index": [0, 2], "expected": 0.9 * 4.0 - 3.0, "atol": 1e-3},
        {"kind": "adv_at", "index": [0, 5], "expected": 0.9 * 1.0 - 2.0, "atol": 1e-3},
    ],
})
cases.append({
    "name": "simple_values_interface_with_truncation", "ref": "multistep_test.py:364-416",
    "inputs": {"r_t": [[1.0, 0.0, 0.0, 0.5, 0.0, 0.0]], "discount_t": [[0.9] * 6], "lambda_": 1.0,
               "values": [[1.0, 2.0, 3.0, 1.0, 1.5, 2.0, 100.0]],
               "truncation_t": [[0.0, 0.0, 1.0, 0.0, 0.0, 1.0]]},
    "checks": [
        {"kind": "adv_at", "index": [0, 2], "expected": 0.9 * 1.0 - 3.0, "atol": 1e-3},
        {"kind": "adv_at", "index": [0, 5], "expected": 0.9 * 100 - 2, "atol": 1e-3},
    ],
})
# "remove truncation and use termination in the same position": the two interfaces then agree (:422-429)
cases.append({
    "name": "termination_makes_interfaces_agree", "ref": "multistep_test.py:420-429",
    "inputs": {"r_t": [[1.0, 0.0, 0.0, 0.5, 0.0, 0.0]], "discount_t": [[1.0, 1.0, 0.0, 1.0, 1.0, 0.0]],
               "lambda_": 1.0, "v_tm1": [[1.0, 2.0, 3.0, 1.0, 1.5, 2.0]], "v_t": [[2.0, 3.0, 4.0, 1.5, 2.0, 1.0]]},
    "checks": [{"kind": "same_as_values", "values": [[1.0, 2.0, 3.0, 1.0, 1.5, 2.0, 100.0]], "atol": 1e-6}],
})

# autoreset with different initial values, multistep_test.py:431-473
cases.append({
    "name": "autoreset_with_initial_values", "ref": "multistep_test.py:431-473",
    "inputs": {"r_t": [[0.0, 0.0, 1.0, 0.0, 0.0]], "discount_t": [[0.9] * 5], "lambda_": 1.0,
               "v_tm1": [[5.0, 4.0, 3.0, 1.0, 2.0]], "v_t": [[4.0, 3.0, 1.0, 2.0, 0.0]],
               "truncation_t": [[0.0, 0.0, 1.0, 0.0, 0.0]]},
    "checks": [
        {"kind": "adv_at", "index": [0, 2], "expected": 1.0 + 0.9 * 1.0 - 3.0, "atol": 1e-3},
        {"kind": "adv_at", "index": [0, 3], "expected": -1.0, "atol": 1e-3},
    ],
})

# all truncated, multistep_test.py:477-497: every advantage is its own TD error
cases.append({
    "name": "all_truncated", "ref": "multistep_test.py:477-497",
    "inputs": {"r_t": [[1.0, 0.5, -0.5]], "discount_t": [[0.9, 0.9, 0.9]], "lambda_": 1.0,
               "values": [[1.0, 2.0, 1.5, 1.0]], "truncation_t": [[1.0, 1.0, 1.0]]},
    "checks": [
        {"kind": "adv_at", "index": [0, 0], "expected": 1.0 + 0.9 * 2.0 - 1.0, "atol": 5e-4},
        {"kind": "adv_at", "index": [0, 1], "expected": 0.5 + 0.9 * 1.5 - 2.0, "atol": 5e-4},
        {"kind": "adv_at", "index": [0, 2], "expected": -0.5 + 0.9 * 1.0 - 1.5, "atol": 5e-4},
    ],
})

# time-major equivalence, multistep_test.py:499-525
cases.append({
    "name": "time_major_equivalence", "ref": "multistep_test.py:499-525",
    "inputs": {"r_t": r_t, "discount_t": discount_t, "lambda_": 1.0, "values": values},
    "checks": [{"kind": "same_time_major", "atol": 1e-6}],
})

out = Path(__file__).with_name("multistep_vectors.json")
out.write_text(json.dumps({"source": "stoix/tests/multistep_test.py @ reference 8fff19c", "cases": cases}, indent=1))
print(f"wrote {out} ({len(cases)} cases)")
