"""CPU regeneration of the synthetic Box environment (stoix_b200/csrc/stx_env.cu) -- TEST
INFRASTRUCTURE.  Philox4x32-10 in NumPy with the same counter/key layout as the kernel, so the
tests can check the kernel's draws, flags and episode bookkeeping element by element.  The env
contract itself (auto-reset with next_obs in extras, episode metrics) follows the reference's call
sites: stoix/systems/ppo/anakin/ff_ppo.py:104-116, stoix/utils/make_env.py:56-60,
stoix/wrappers/envpool.py:94-133."""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
TAG_OBS, TAG_RESET, TAG_STEP = 0x4F425331, 0x52535431, 0x53545031


def philox4x32(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(c, np.uint64) & 0xFFFFFFFF for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0, k1 = np.uint64(k0), np.uint64(k1)
    for _ in range(10):
        p0 = np.uint64(M0) * c0
        p1 = np.uint64(M1) * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & np.uint64(0xFFFFFFFF)
        hi1, lo1 = p1 >> np.uint64(32), p1 & np.uint64(0xFFFFFFFF)
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0), lo1, (hi0 ^ c3 ^ k1), lo0
        k0 = (k0 + np.uint64(W0)) & np.uint64(0xFFFFFFFF)
        k1 = (k1 + np.uint64(W1)) & np.uint64(0xFFFFFFFF)
    return c0, c1, c2, c3


def u01(x):
    # 23 bits + half-step offset (stx_common.cuh u01): k + 0.5 is exact in fp32, the result is never 0 or 1
    return ((np.asarray(x, np.uint64) >> np.uint64(9)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 8388608.0)


def normal2(a, b):
    r = np.sqrt(np.float32(-2.0) * np.log(u01(a))).astype(np.float32)
    ang = (np.float32(2.0) * u01(b)).astype(np.float64) * np.pi
    return (r * np.cos(ang)).astype(np.float32), (r * np.sin(ang)).astype(np.float32)


class SynthEnvOracle:
    def __init__(self, E, D, seed, p_term, p_trunc):
        self.E, self.D, self.seed = E, D, seed
        self.p_term, self.p_trunc = np.float32(p_term), np.float32(p_trunc)
        self.run_return = np.zeros(E, np.float32)
        self.run_length = np.zeros(E, np.int32)

    def _obs(self, step, tag):
        E, D = self.E, self.D
        groups = (D + 3) // 4
        e = np.arange(E, dtype=np.uint64)[:, None]
        g = np.arange(groups, dtype=np.uint64)[None, :]
        hi = (((e >> np.uint64(32)) << np.uint64(16)) ^ np.uint64(step >> 32)) & np.uint64(0xFFFFFFFF)
        r = philox4x32(e, np.uint64(step & 0xFFFFFFFF), hi ^ (g << np.uint64(8)), np.uint64(tag), self.seed & 0xFFFFFFFF, self.seed >> 32)
        a0, a1 = normal2(r[0], r[1])
        b0, b1 = normal2(r[2], r[3])
        out = np.stack([a0, a1, b0, b1], axis=-1).reshape(E, groups * 4)
        return out[:, :D]

    def step(self, step):
        E = self.E
        e = np.arange(E, dtype=np.uint64)
        hi = (((e >> np.uint64(32)) << np.uint64(16)) ^ np.uint64(step >> 32)) & np.uint64(0xFFFFFFFF)
        r = philox4x32(e, np.uint64(step & 0xFFFFFFFF), hi, np.uint64(TAG_STEP), self.seed & 0xFFFFFFFF, self.seed >> 32)
        reward, _ = normal2(r[0], r[1])
        term = u01(r[2]) < self.p_term
        trunc = (~term) & (u01(r[3]) < self.p_trunc)
        last = term | trunc
        nxt = self._obs(step, TAG_OBS)
        rst = self._obs(step, TAG_RESET)
        obs = np.where(last[:, None], rst, nxt)
        ret = self.run_return + reward
        ln = self.run_length + 1
        self.run_return = np.where(last, np.float32(0), ret).astype(np.float32)
        self.run_length = np.where(last, 0, ln).astype(np.int32)
        return dict(reward=reward, done=term.astype(np.uint8), truncated=trunc.astype(np.uint8),
                    is_terminal=last.astype(np.uint8), next_obs=nxt, obs=obs, ep_return=ret, ep_length=ln)
