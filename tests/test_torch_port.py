"""The torch-CPU port timed as the CPU baseline must compute the same update as the NumPy oracle."""
import numpy as np
import torch

from oracle import ppo_oracle as O
from oracle.torch_cpu_ppo import CpuAnakinPPO


def test_port_update_matches_numpy_oracle():
    torch.manual_seed(0)
    m = CpuAnakinPPO(E=16, T=8, D=6, A=3, hidden=(16, 16), epochs=2, num_minibatches=2, num_updates=5, p_term=0.1, p_trunc=0.1)
    for p in m.actor + m.critic:
        p.add_(torch.randn(p.shape, generator=m.gen) * 0.1)
    tr = m.rollout()
    to_o = lambda ps: O.MLPParams([ps[i].double().numpy().copy() for i in range(0, len(ps), 2)], [ps[i].double().numpy().copy() for i in range(1, len(ps), 2)])
    actor, critic = to_o(m.actor), to_o(m.critic)
    traj = O.Trajectory(obs=tr["obs"].double().numpy(), action=tr["action"].numpy(), reward=tr["reward"].double().numpy(),
                        done=tr["done"].numpy(), truncated=tr["trunc"].numpy(), next_obs=tr["obs"].double().numpy(),
                        value=tr["value"].double().numpy(), bootstrap_value=tr["bootstrap"].double().numpy(),
                        log_prob=tr["log_prob"].double().numpy())
    rng = np.random.default_rng(0)
    perms = np.stack([rng.permutation(128) for _ in range(2)])
    h = O.PPOHyper(epochs=2, num_minibatches=2, num_updates=5)
    a_st = O.AdamState(np.zeros(actor.flat().size), np.zeros(actor.flat().size))
    c_st = O.AdamState(np.zeros(critic.flat().size), np.zeros(critic.flat().size))
    a2, c2, metrics, adv, tgt = O.ppo_update(actor, critic, a_st, c_st, traj, perms, h)
    info = m.update(tr, perms)
    np.testing.assert_allclose(to_o(m.actor).flat(), a2.flat(), rtol=2e-4, atol=2e-6)
    np.testing.assert_allclose(to_o(m.critic).flat(), c2.flat(), rtol=2e-4, atol=2e-6)
    np.testing.assert_allclose(info["value_loss"], metrics["value_loss"][-1, -1], rtol=1e-4)
    adv_t, tgt_t = m.gae(tr["reward"], tr["value"], tr["bootstrap"], tr["done"], tr["trunc"], 0.99, 0.95)
    np.testing.assert_allclose(adv_t.numpy(), adv, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(tgt_t.numpy(), tgt, rtol=1e-4, atol=1e-5)
