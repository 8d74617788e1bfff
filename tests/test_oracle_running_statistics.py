"""Self-consistency of the oracle for the observation-normalisation branch (oracle/running_statistics.py; the reference
has no test for it): the batched Welford update must reproduce plain batch statistics, be independent of how the data is
split into updates and replicas (the psum), and keep the order `normalise with the old statistics, then update`."""
import numpy as np
import pytest

from oracle import running_statistics as RS


def test_updates_reproduce_batch_mean_and_population_std():
    rng = np.random.default_rng(0)
    D = 7
    data = rng.standard_normal((5, 16, 4, D)) * rng.uniform(0.1, 5.0, D) + rng.uniform(-3, 3, D)  # 5 updates of (T=16, E=4)
    st = RS.initialize((D,))
    assert np.array_equal(RS.normalize(data[0], st), data[0])  # initial state: mean 0, std 1 -> identity
    for k in range(5):
        st = RS.update(st, [data[k]])
        seen = data[: k + 1].reshape(-1, D)
        assert st.count == seen.shape[0]
        np.testing.assert_allclose(st.mean, seen.mean(0), rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(st.std, seen.std(0), rtol=1e-10)  # summed_variance / count = population variance


@pytest.mark.parametrize("world", [2, 4])
def test_the_psum_over_replicas_equals_one_update_on_the_concatenation(world):
    rng = np.random.default_rng(world)
    D = 5
    first = rng.standard_normal((8, 3, D))
    shards = [rng.standard_normal((8, 3, D)) * (r + 1) + r for r in range(world)]
    a = RS.update(RS.update(RS.initialize((D,)), [first]), shards)
    b = RS.update(RS.update(RS.initialize((D,)), [first]), [np.concatenate(shards, axis=1)])
    assert a.count == b.count
    np.testing.assert_allclose(a.mean, b.mean, rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(a.summed_variance, b.summed_variance, rtol=1e-11)
    np.testing.assert_allclose(a.std, b.std, rtol=1e-12)


def test_std_limits_and_integer_passthrough():
    st = RS.update(RS.initialize((3,)), [np.tile(np.array([1.0, -2.0, 0.0]), (10, 1))], std_min_value=5e-4, std_max_value=5e4)
    np.testing.assert_array_equal(st.std, [5e-4] * 3)  # a constant feature is clamped to the minimum, not divided by zero
    big = RS.update(RS.initialize((1,)), [np.array([[-1e9], [1e9]])], std_min_value=5e-4, std_max_value=5e4)
    assert big.std[0] == 5e4
    ints = np.arange(6).reshape(2, 3)
    assert RS.normalize(ints, st) is not None and np.array_equal(RS.normalize(ints, st), ints)  # only inexact dtypes are touched
    assert np.abs(RS.normalize(np.array([[1e9, 0.0, 0.0]]), st, max_abs_value=10.0)).max() == 10.0


def test_update_step_normalises_with_the_statistics_from_before_the_update():
    rng = np.random.default_rng(5)
    D = 4
    st0 = RS.update(RS.initialize((D,)), [rng.standard_normal((32, D)) * 2 + 1])
    shards = [rng.standard_normal((6, 2, D)) + 3 for _ in range(2)]
    normed, st1 = RS.ppo_update_step_statistics(st0, shards)
    for n, o in zip(normed, shards):
        np.testing.assert_allclose(n, (o - st0.mean) / st0.std, rtol=1e-13)
    assert st1.count == st0.count + 2 * 12 and not np.allclose(st1.mean, st0.mean)
