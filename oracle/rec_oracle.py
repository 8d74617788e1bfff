"""CPU restatement of the recurrent PPO path: ScannedRNN(GRU) with episode resets, RecurrentActor /
RecurrentCritic, and one `_update_step` of stoix/systems/ppo/anakin/rec_ppo.py.  TEST INFRASTRUCTURE ONLY -- never imported by
the product (stoix_b200/).

PARITY UNPINNED: the reference has no tests or golden vectors for this path and JAX / Flax cannot run here; the restatement
follows the call sites (file:line below) and the published definition of flax.linen.GRUCell (the pinned flax is not vendored):
    r = sigmoid(W_ir x + b_ir + W_hr h)            dense_i: use_bias=True,  dense_h: use_bias=False
    z = sigmoid(W_iz x + b_iz + W_hz h)
    n = tanh(W_in x + b_in + r * (W_hn h + b_hn))  ("add bias because the linear transformations aren't directly summed")
    h' = (1 - z) * n + z * h
and is cross-checked against torch.autograd on CPU (tests/test_oracle_rec.py).

Reference behaviour restated literally, including what looks unusual:
  * the hidden state is reset to zeros where the PREVIOUS transition ended (done | truncated), networks/base.py:139-148;
  * the transition stores the hidden state AFTER its step (rec_ppo.py:118-129) and the loss re-runs the network from
    `hstates[0]` of the chunk (rec_ppo.py:216-221, 243-247), i.e. from the state after the first step;
  * GAE uses discount_t = (1 - last_done_t) * gamma (the done flag stored with the transition is the one BEFORE the step,
    rec_ppo.py:170-182) through the `values=` interface, no truncation argument; the bootstrap value is masked by last_done;
  * minibatches are column subsets of the batch reshaped to (chunk, num_envs * num_chunks) by a plain reshape (rec_ppo.py:327-352).
Parameter layout of one network (flat, fp32): [pre-torso Dense layers | W_i (P x 3H, columns r|z|n) | b_i (3H) | W_h (H x 3H) | b_hn (H) |
post-torso Dense layers | head]: `pre` = MLP(D, *pre_sizes, 3H) whose "head" is the input projection of the cell, `post` = MLP(H, *post_sizes, A)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Tuple

import numpy as np

from oracle import ppo_oracle as O
from oracle.sac_oracle import q_input_grad as mlp_input_grad


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


@dataclass
class RecNet:
    pre: O.MLPParams      # (D, *pre_sizes, 3H)
    Wh: np.ndarray        # (H, 3H)
    bhn: np.ndarray       # (H,)
    post: O.MLPParams     # (H, *post_sizes, out)

    @property
    def H(self) -> int:
        return self.Wh.shape[0]

    @property
    def is_lstm(self) -> bool:   # W_h (H, 4H), no separate b_hn (empty): the four hidden biases are the bias of the input projection
        return self.Wh.shape[1] == 4 * self.Wh.shape[0]

    def flat(self) -> np.ndarray:
        return np.concatenate([self.pre.flat(), self.Wh.ravel(), self.bhn.ravel(), self.post.flat()])

    @staticmethod
    def from_flat(flat: np.ndarray, pre_sizes, H: int, post_sizes, activation: str = "silu") -> "RecNet":
        n_pre = sum(pre_sizes[i] * pre_sizes[i + 1] + pre_sizes[i + 1] for i in range(len(pre_sizes) - 1))
        o = n_pre
        pre = O.MLPParams.from_flat(flat[:o], list(pre_sizes), activation)
        G = pre_sizes[-1] // H          # 3 gate blocks: GRU (followed by b_hn), 4: LSTM (no extra bias)
        Wh = flat[o:o + H * G * H].reshape(H, G * H).copy()
        o += H * G * H
        nb = H if G == 3 else 0
        bhn = flat[o:o + nb].copy()
        o += nb
        post = O.MLPParams.from_flat(flat[o:], list(post_sizes), activation)
        return RecNet(pre, Wh, bhn, post)


def gru_forward(gi: np.ndarray, reset: np.ndarray, h0: np.ndarray, Wh: np.ndarray, bhn: np.ndarray):
    """ScannedRNN over T steps (networks/base.py:124-152).  gi (T, E, 3H) = input projections incl. their biases, reset (T, E) bool."""
    T, E, H3 = gi.shape
    H = H3 // 3
    h = h0
    hs, cache = [], []
    for t in range(T):
        hp = np.where(reset[t][:, None], 0.0, h)          # base.py:139-148: reset BEFORE the cell
        gh = hp @ Wh
        ghn = gh[:, 2 * H:] + bhn
        r = _sigmoid(gi[t][:, :H] + gh[:, :H])
        z = _sigmoid(gi[t][:, H:2 * H] + gh[:, H:2 * H])
        n = np.tanh(gi[t][:, 2 * H:] + r * ghn)
        h = (1.0 - z) * n + z * hp
        hs.append(h)
        cache.append((hp, r, z, n, ghn))
    return np.stack(hs), cache


def gru_backward(cache, reset: np.ndarray, d_hseq: np.ndarray, Wh: np.ndarray):
    """Reverse-mode of gru_forward: d(loss)/d(gi), d(Wh), d(bhn), d(h0) given d(loss)/d(h_t) for every t."""
    T, E, H = d_hseq.shape
    d_gi = np.zeros((T, E, 3 * H), d_hseq.dtype)
    dWh = np.zeros_like(Wh)
    dbhn = np.zeros(H, d_hseq.dtype)
    dh_rec = np.zeros((E, H), d_hseq.dtype)
    for t in range(T - 1, -1, -1):
        hp, r, z, n, ghn = cache[t]
        dh = d_hseq[t] + dh_rec
        dn, dz, dhp = dh * (1.0 - z), dh * (hp - n), dh * z
        dpre_n = dn * (1.0 - n * n)
        dr = dpre_n * ghn
        dpre_r, dpre_z = dr * r * (1.0 - r), dz * z * (1.0 - z)
        dgh = np.concatenate([dpre_r, dpre_z, dpre_n * r], axis=1)
        d_gi[t] = np.concatenate([dpre_r, dpre_z, dpre_n], axis=1)
        dWh += hp.T @ dgh
        dbhn += dgh[:, 2 * H:].sum(0)
        dhp = dhp + dgh @ Wh.T
        dh_rec = np.where(reset[t][:, None], 0.0, dhp)
    return d_gi, dWh, dbhn, dh_rec


def lstm_forward(gi: np.ndarray, reset: np.ndarray, carry0: np.ndarray, Wh: np.ndarray):
    """ScannedRNN(cell_type="lstm") over T steps.  flax.linen.LSTMCell (published definition; the hidden Denses carry the biases):
        i = sigmoid(W_ii x + W_hi h + b_hi)   f = sigmoid(W_if x + W_hf h + b_hf)   g = tanh(W_ig x + W_hg h + b_hg)
        o = sigmoid(W_io x + W_ho h + b_ho)   c' = f c + i g   h' = o tanh(c')      carry = (c', h'), output = h'
    gi (T, E, 4H) = W_i x + b_h (columns i|f|g|o: the four hidden biases ride on the input projection, the sum is the same),
    carry0 (E, 2H) = (c | h), Wh (H, 4H).  Both halves of the carry are zeroed where reset_t is set (base.py:139-148 tree_map)."""
    T, E, H4 = gi.shape
    H = H4 // 4
    c, h = carry0[:, :H], carry0[:, H:]
    hs, cs, cache = [], [], []
    for t in range(T):
        cp = np.where(reset[t][:, None], 0.0, c)
        hp = np.where(reset[t][:, None], 0.0, h)
        z = gi[t] + hp @ Wh
        i, f, g, o = _sigmoid(z[:, :H]), _sigmoid(z[:, H:2 * H]), np.tanh(z[:, 2 * H:3 * H]), _sigmoid(z[:, 3 * H:])
        c = f * cp + i * g
        tc = np.tanh(c)
        h = o * tc
        hs.append(h), cs.append(c)
        cache.append((cp, hp, i, f, g, o, tc))
    return np.stack(hs), np.concatenate([c, h], 1), cache


def lstm_backward(cache, reset: np.ndarray, d_hseq: np.ndarray, Wh: np.ndarray):
    """Reverse-mode of lstm_forward: d(gi), d(Wh), d(carry0) given d(loss)/d(h_t) for every t."""
    T, E, H = d_hseq.shape
    d_gi = np.zeros((T, E, 4 * H), d_hseq.dtype)
    dWh = np.zeros_like(Wh)
    dh_rec, dc_rec = np.zeros((E, H), d_hseq.dtype), np.zeros((E, H), d_hseq.dtype)
    for t in range(T - 1, -1, -1):
        cp, hp, i, f, g, o, tc = cache[t]
        dh = d_hseq[t] + dh_rec
        do = dh * tc
        dc = dc_rec + dh * o * (1.0 - tc * tc)
        di, df, dg, dcp = dc * g, dc * cp, dc * i, dc * f
        dz = np.concatenate([di * i * (1 - i), df * f * (1 - f), dg * (1 - g * g), do * o * (1 - o)], axis=1)
        d_gi[t] = dz
        dWh += hp.T @ dz
        dhp = dz @ Wh.T
        dh_rec = np.where(reset[t][:, None], 0.0, dhp)
        dc_rec = np.where(reset[t][:, None], 0.0, dcp)
    return d_gi, dWh, np.concatenate([dc_rec, dh_rec], 1)


def rec_forward(net: RecNet, h0: np.ndarray, obs: np.ndarray, reset: np.ndarray):
    """RecurrentActor / RecurrentCritic.__call__ (networks/base.py:162-222): obs (T, E, D) -> out (T, E, A), last hidden state."""
    T, E, D = obs.shape
    gi, c_pre = O.mlp_forward(net.pre, obs.reshape(T * E, D))
    if net.is_lstm:   # carry = (c | h), (E, 2H)
        h_seq, last, c_gru = lstm_forward(gi.reshape(T, E, -1), reset, h0, net.Wh)
    else:
        h_seq, c_gru = gru_forward(gi.reshape(T, E, -1), reset, h0, net.Wh, net.bhn)
        last = h_seq[-1]
    out, c_post = O.mlp_forward(net.post, h_seq.reshape(T * E, -1))
    return out.reshape(T, E, -1), last, (c_pre, c_gru, c_post, h_seq, reset)


def rec_backward(net: RecNet, cache, d_out: np.ndarray) -> RecNet:
    c_pre, c_gru, c_post, h_seq, reset = cache
    T, E, H = h_seq.shape
    d2 = d_out.reshape(T * E, -1)
    g_post = O.mlp_backward(net.post, c_post, d2)
    d_h = mlp_input_grad(net.post, c_post, d2).reshape(T, E, H)
    if net.is_lstm:
        d_gi, dWh, _ = lstm_backward(c_gru, reset, d_h, net.Wh)
        dbhn = np.zeros_like(net.bhn)
    else:
        d_gi, dWh, dbhn, _ = gru_backward(c_gru, reset, d_h, net.Wh)
    g_pre = O.mlp_backward(net.pre, c_pre, d_gi.reshape(T * E, -1))
    return RecNet(g_pre, dWh, dbhn, g_post)


@dataclass
class RecTrajectory:
    """RNNPPOTransition fields, time-major (rec_ppo.py:118-129): done / truncated are the flags BEFORE the step."""
    obs: np.ndarray          # (T, E, D)
    done: np.ndarray         # (T, E) bool  last_done
    truncated: np.ndarray    # (T, E) bool  last_truncated
    action: np.ndarray       # (T, E)
    value: np.ndarray        # (T, E)
    reward: np.ndarray       # (T, E)
    log_prob: np.ndarray     # (T, E)
    h_actor: np.ndarray      # (T, E, H)  hidden state AFTER the step
    h_critic: np.ndarray     # (T, E, H)
    last_val: np.ndarray     # (E,) critic(last obs) masked by last_done (rec_ppo.py:160-163)


def rec_gae(traj: RecTrajectory, gamma: float, lam: float, standardize: bool = True):
    """rec_ppo.py:165-179."""
    v_t = np.concatenate([traj.value, traj.last_val[None]], 0)
    d_t = (1.0 - traj.done.astype(np.float64)) * gamma
    return O.gae(traj.reward, d_t, lam, values=v_t, time_major=True, standardize_advantages=standardize)


def rec_minibatch_grads(actor: RecNet, critic: RecNet, traj: RecTrajectory, adv, tgt, cols: np.ndarray, chunk: int, h: O.PPOHyper):
    """`_update_minibatch` up to the gradients (rec_ppo.py:196-262) on the columns `cols` of the (chunk, E * num_chunks) batch."""
    T, E = traj.reward.shape
    nc = T // chunk
    r2 = lambda x: x.reshape((chunk, E * nc) + x.shape[2:])[:, cols]
    obs, action, logp_old, v_old = r2(traj.obs), r2(traj.action), r2(traj.log_prob), r2(traj.value)
    reset = r2(traj.done) | r2(traj.truncated)
    a_, t_ = r2(adv), r2(tgt)
    m = chunk * cols.size
    # actor (rec_ppo.py:207-231)
    logits, _, a_cache = rec_forward(actor, r2(traj.h_actor)[0], obs, reset)
    _, dlg, a_info = O.actor_loss_and_dlogits(logits.reshape(m, -1), action.reshape(m), logp_old.reshape(m), a_.reshape(m), h.clip_eps, h.ent_coef)
    ga = rec_backward(actor, a_cache, dlg.reshape(chunk, cols.size, -1))
    # critic (rec_ppo.py:233-257)
    val, _, c_cache = rec_forward(critic, r2(traj.h_critic)[0], obs, reset)
    _, dv, c_info = O.critic_loss_and_dvalue(val.reshape(m), v_old.reshape(m), t_.reshape(m), h.clip_eps, h.vf_coef)
    gc = rec_backward(critic, c_cache, dv.reshape(chunk, cols.size, 1))
    return ga, gc, {**a_info, **c_info}
