"""Executable model of the mbarrier protocol of K3a (`tc_ppo_fwd_bwd_kernel`, stoix_b200/csrc/stx_tc_ppo.cu) under random
schedules.  The kernel splits every 256-wide GEMM into four 64-column parts that trail the running epilogue part by part;
the hand-off uses phase-parity waits on `d_ready[4]`, `chunk_done[4]`, `head_ready`, `head_done`, `x_full[2]`, `x_empty[2]`.
This model restates the three warp roles with the SAME wait/arrive sequence and parities as the kernel, adds an
asynchronous in-order tensor pipe (tcgen05.mma completes later than it is issued; tcgen05.commit arrives on an mbarrier
when everything issued before it has completed) and checks, for many random interleavings:
  * no deadlock, and no waiter ever falls two phases behind a barrier (the ABA hazard of parity waits);
  * every tensor-memory region holds what its reader expects (D parts, the h1 / h2|dh2 operand regions, the X stages):
    nothing is overwritten before its last reader is done and nothing is read before its producer has completed.
It is test infrastructure for the protocol only -- it does no arithmetic."""
import random

import pytest

PARTS = 4


class MBar:
    def __init__(self, count, name):
        self.count, self.pending, self.phase, self.name = count, count, 0, name

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, f"{self.name}: more arrivals than the barrier expects in one phase"
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def done(self, parity):  # mbarrier.try_wait.parity: true once the phase of that parity has completed
        return (self.phase & 1) != parity


class Model:
    def __init__(self, tiles, rng, mutation=None):
        self.rng, self.tiles, self.mut = rng, tiles, mutation  # `mutation`: a deliberately broken variant (the model must catch it)
        self.d_ready = [MBar(1, f"d_ready{p}") for p in range(PARTS)]
        self.chunk_done = [MBar(2, f"chunk_done{p}") for p in range(PARTS)]  # kernel: 8 = 4 lane quarters x 2 halves; model: 2 halves
        self.head_ready, self.head_done = MBar(1, "head_ready"), MBar(1, "head_done")  # kernel head_done: 4 quarters; model: 1
        self.x_full, self.x_empty = [MBar(1, "x_full0"), MBar(1, "x_full1")], [MBar(1, "x_empty0"), MBar(1, "x_empty1")]
        # contents: what a region currently holds, as (kind, tile); readers assert on it
        self.D = [None] * PARTS          # accumulator parts: ("G0".."G4", tile) once COMPLETE, ("busy", ...) while MMAs are in flight
        self.head = None                 # D columns 0..15 after G2
        self.A1 = [None] * 8             # per 32-column chunk: ("h1", tile)
        self.A2 = [None] * 8             # ("h2", tile) or ("dh2", tile)
        self.X = [None, None]            # smem stages
        self.dz = None
        self.pipe = []                   # in-order tensor pipe: issued, not yet executed operations
        self.waits = {}                  # (actor, barrier name) -> last phase index waited for (ABA check)

    # ---- helpers used by the actors (generators) ----
    def wait(self, who, bar, parity):
        while not bar.done(parity):
            yield
        # the phase this waiter consumed must be the most recent completed one: never two behind
        completed = bar.phase - 1
        key = (who, bar.name)
        prev = self.waits.get(key, -1)
        assert completed == prev + 1, f"{who} consumed phase {completed} of {bar.name} after phase {prev}: skipped or repeated a phase"
        self.waits[key] = completed

    def pipe_step(self):
        """Execute the oldest issued tensor-pipe operation (called by the scheduler at random times)."""
        if not self.pipe:
            return False
        op = self.pipe.pop(0)
        op()
        return True


def producer(m):
    """Gather warps: X tile `it` into stage it&1 (waits x_empty for it >= 2)."""
    for it in range(m.tiles):
        s = it & 1
        if it >= 2 and m.mut != "producer ignores x_empty":
            yield from m.wait("prod", m.x_empty[s], ((it >> 1) & 1) ^ 1)
        assert m.X[s] is None or m.X[s][1] == it - 2, f"X stage {s} overwritten while it holds {m.X[s]}"
        m.X[s] = ("x", it)
        yield
        m.x_full[s].arrive()


def mma_warp(m):
    def issue_mma(check, effect):
        def op():
            check()
            effect()
        m.pipe.append(op)

    def commit(bar):
        m.pipe.append(bar.arrive)

    for it in range(m.tiles):
        s = it & 1
        yield from m.wait("mma", m.x_full[s], (it >> 1) & 1)
        # ---- G0 (trailing E4 of the previous tile) ----
        for pt in range(PARTS):
            if it > 0 and m.mut != "G0 does not wait for E4":
                yield from m.wait("mma", m.chunk_done[pt], 1)

            def chk(pt=pt, it=it, s=s):
                assert m.X[s] == ("x", it), f"G0 tile {it} reads X stage {s} = {m.X[s]}"
                assert m.D[pt] is None or m.D[pt] == ("free", it - 1), f"G0 tile {it} overwrites D[{pt}] = {m.D[pt]}"
            issue_mma(chk, lambda pt=pt, it=it: m.D.__setitem__(pt, ("G0", it)))
            commit(m.d_ready[pt])
            if pt == 3:
                def release(s=s, it=it):
                    m.x_empty[s].arrive()
                m.pipe.append(release)
            yield
        # ---- G1 (trailing E0) and G4 (trailing E3): parts <= j, K groups as in the kernel ----
        def trailing(gemm, parity, src, src_kind, prev):
            for j in range(PARTS):
                yield from m.wait("mma", m.chunk_done[j], parity)
                for pt in range(j + 1):
                    for g in (range(0, j + 1) if pt == j else [j]):
                        def chk(pt=pt, g=g, it=it):
                            for c in (2 * g, 2 * g + 1):  # K group g = operand chunks 2g, 2g+1
                                assert src[c] == (src_kind, it), f"{gemm} tile {it} part {pt} reads {src_kind} chunk {c} = {src[c]}"
                            if g == 0:
                                assert m.D[pt] == ("free", it) and True, f"{gemm} tile {it} overwrites D[{pt}] = {m.D[pt]}"
                        def eff(pt=pt, g=g, it=it):
                            m.D[pt] = ("busy", gemm, it, g)
                        issue_mma(chk, eff)
                    if j == 3:
                        def fin(pt=pt, it=it):
                            assert m.D[pt] == ("busy", gemm, it, 3)
                            m.D[pt] = (gemm, it)
                        m.pipe.append(fin)
                        commit(m.d_ready[pt])
                yield
        yield from trailing("G1", 0, m.A1, "h1", "G0")
        # ---- G2 (head, trailing E1) ----
        for j in range(PARTS):
            yield from m.wait("mma", m.chunk_done[j], 1)
            def chk(j=j, it=it):
                for c in (2 * j, 2 * j + 1):
                    assert m.A2[c] == ("h2", it), f"G2 tile {it} reads h2 chunk {c} = {m.A2[c]}"
                assert m.D[0] == ("free", it), f"G2 tile {it} writes the head columns while D[0] = {m.D[0]}"
            issue_mma(chk, lambda: None)
            if j == 3:
                m.pipe.append(lambda it=it: setattr(m, "head", ("G2", it)))
                commit(m.head_ready)
            yield
        # ---- G3 (after E2) ----
        if m.mut != "G3 does not wait for head_done":
            yield from m.wait("mma", m.head_done, it & 1)
        def chk3(it=it):
            assert m.dz == ("dz", it)
            for pt in range(PARTS):
                assert m.D[pt] == ("free", it), f"G3 tile {it} overwrites D[{pt}] = {m.D[pt]}"
        def eff3(it=it):
            for pt in range(PARTS):
                m.D[pt] = ("G3", it)
        issue_mma(chk3, eff3)
        for pt in range(PARTS):
            commit(m.d_ready[pt])
        yield
        yield from trailing("G4", 1 if m.mut == "G4 waits with E1's parity" else 0, m.A2, "dh2", "G3")


def epilogue(m, half):
    who = f"epi{half}"
    for it in range(m.tiles):
        # ---- E0 / E1 ----
        for layer, (gemm, dst, kind) in enumerate((("G0", m.A1, "h1"), ("G1", m.A2, "h2"))):
            for cc in range(PARTS):
                c = cc * 2 + half
                yield from m.wait(who, m.d_ready[cc], 0 if m.mut == "E1 waits with E0's parity" else layer)
                assert m.D[cc] == (gemm, it), f"E{layer} tile {it} reads D[{cc}] = {m.D[cc]}, wants {gemm}"
                yield
                if kind == "h1":
                    assert dst[c] is None or dst[c] == ("h1", it - 1), f"E0 tile {it} overwrites A1[{c}] = {dst[c]}"
                else:
                    assert dst[c] is None or dst[c] == ("dh2", it - 1), f"E1 tile {it} overwrites A2[{c}] = {dst[c]}"
                    # the previous tile's G4 must have finished reading dh2: its parts were all consumed by E4 already
                dst[c] = (kind, it)
                if m.chunk_done[cc].pending == 1:  # second arrival of this part: both chunk halves have left D
                    m.D[cc] = ("free", it)
                m.chunk_done[cc].arrive()
                yield
        # ---- E2 (half 0 only) ----
        if half == 0:
            yield from m.wait(who, m.head_ready, it & 1)
            assert m.head == ("G2", it), f"E2 tile {it} reads head = {m.head}"
            yield
            m.dz = ("dz", it)
            m.head_done.arrive()
        # ---- E3 / E4 ----
        for layer, (gemm, par) in ((1, ("G3", 0)), (0, ("G4", 1))):
            for cc in range(PARTS):
                c = cc * 2 + half
                yield from m.wait(who, m.d_ready[cc], par)
                assert m.D[cc] == (gemm, it), f"E{4 - layer} tile {it} reads D[{cc}] = {m.D[cc]}, wants {gemm}"
                if layer == 1:
                    assert m.A2[c] == ("h2", it), f"E3 tile {it} reads the relu mask A2[{c}] = {m.A2[c]}"
                    yield
                    m.A2[c] = ("dh2", it)  # G2 (reader of h2) completed before head_ready, which this warp's G3 wait implies
                else:
                    assert m.A1[c] == ("h1", it), f"E4 tile {it} reads the relu mask A1[{c}] = {m.A1[c]}"
                    yield
                if m.chunk_done[cc].pending == 1:
                    m.D[cc] = ("free", it)
                m.chunk_done[cc].arrive()
                yield


def run(seed, tiles, mutation=None):
    rng = random.Random(seed)
    m = Model(tiles, rng, mutation)
    actors = {"prod": producer(m), "mma": mma_warp(m), "epi0": epilogue(m, 0), "epi1": epilogue(m, 1)}
    idle = 0
    while actors or m.pipe:
        choices = list(actors) + (["pipe"] if m.pipe else [])
        pick = rng.choice(choices)
        before = (tuple(b.phase for b in m.d_ready + m.chunk_done + [m.head_ready, m.head_done] + m.x_full + m.x_empty), len(m.pipe))
        if pick == "pipe":
            m.pipe_step()
        else:
            try:
                next(actors[pick])
            except StopIteration:
                del actors[pick]
        after = (tuple(b.phase for b in m.d_ready + m.chunk_done + [m.head_ready, m.head_done] + m.x_full + m.x_empty), len(m.pipe))
        idle = idle + 1 if before == after else 0
        assert idle < 20000, f"deadlock (seed {seed}): still running {sorted(actors)}, pipe {len(m.pipe)}"
    # everything drained: all regions hold the last tile's data
    assert all(d == ("free", tiles - 1) for d in m.D)


@pytest.mark.parametrize("tiles", [1, 2, 4])
def test_k3a_barrier_protocol_under_random_schedules(tiles):
    for seed in range(150):
        run(seed * 7919 + tiles, tiles)


@pytest.mark.parametrize("mutation", ["E1 waits with E0's parity", "G3 does not wait for head_done", "G0 does not wait for E4",
                                      "G4 waits with E1's parity", "producer ignores x_empty"])
def test_the_model_catches_broken_protocols(mutation):
    """Sensitivity of the model: each single-line protocol error is caught (hazard assertion or deadlock) within 150 schedules."""
    with pytest.raises(AssertionError):
        for seed in range(150):
            run(seed * 104729 + 1, 4, mutation)


# ======================================================================================================================
# The forward-only variant of the same scheme: tc_rollout_kernel (stx_tc_rollout.cu) / tc_mlp_fwd_kernel (stx_tc_mlp.cu).
# Per step t: G0 (4 parts, after head_done(t-1)) -> E0 -> G1 (trailing) -> E1 -> G2 (head, trailing) -> E2 reads the head.
# The environment producers run one step ahead: they fill X stage (t+1)&1 during step t (x_empty gate from t >= 1).
# ======================================================================================================================


def ro_producer(m, steps):
    m.X[0] = ("x", 0)  # stage 0 <- the carried observation
    m.x_full[0].arrive()
    yield
    for t in range(steps):
        if t + 1 < steps:
            s = (t + 1) & 1
            if t >= 1:
                yield from m.wait("prod", m.x_empty[s], ((t - 1) >> 1) & 1)
            assert m.X[s] is None or m.X[s][1] == t - 1, f"X stage {s} overwritten while it holds {m.X[s]} (producing step {t + 1})"
            m.X[s] = ("x", t + 1)
            yield
            m.x_full[s].arrive()
        yield


def ro_mma(m, steps):
    def issue(check, effect):
        m.pipe.append(lambda: (check(), effect()))

    for t in range(steps):
        s = t & 1
        yield from m.wait("mma", m.x_full[s], (t >> 1) & 1)
        if t > 0:
            yield from m.wait("mma", m.head_done, (t - 1) & 1)
        for pt in range(PARTS):
            def chk(pt=pt, t=t, s=s):
                assert m.X[s] == ("x", t), f"G0 step {t} reads X stage {s} = {m.X[s]}"
                assert m.D[pt] is None or m.D[pt] == ("free", t - 1), f"G0 step {t} overwrites D[{pt}] = {m.D[pt]}"
                assert pt != 0 or t == 0 or m.head == ("read", t - 1), f"G0 step {t} overwrites the head columns: {m.head}"
            issue(chk, lambda pt=pt, t=t: m.D.__setitem__(pt, ("G0", t)))
            m.pipe.append(m.d_ready[pt].arrive)
        m.pipe.append(m.x_empty[s].arrive)
        yield
        for j in range(PARTS):  # G1 trailing E0
            yield from m.wait("mma", m.chunk_done[j], 0)
            for pt in range(j + 1):
                for g in (range(0, j + 1) if pt == j else [j]):
                    def chk(pt=pt, g=g, t=t):
                        for c in (2 * g, 2 * g + 1):
                            assert m.A1[c] == ("h1", t), f"G1 step {t} reads h1 chunk {c} = {m.A1[c]}"
                        if g == 0:
                            assert m.D[pt] == ("free", t), f"G1 step {t} overwrites D[{pt}] = {m.D[pt]}"
                    issue(chk, lambda pt=pt, g=g, t=t: m.D.__setitem__(pt, ("busy", "G1", t, g)))
                if j == 3:
                    def fin(pt=pt, t=t):
                        assert m.D[pt] == ("busy", "G1", t, 3)
                        m.D[pt] = ("G1", t)
                    m.pipe.append(fin)
                    m.pipe.append(m.d_ready[pt].arrive)
            yield
        for j in range(PARTS):  # head trailing E1
            yield from m.wait("mma", m.chunk_done[j], 1)
            def chk(j=j, t=t):
                for c in (2 * j, 2 * j + 1):
                    assert m.A2[c] == ("h2", t), f"G2 step {t} reads h2 chunk {c} = {m.A2[c]}"
                assert m.D[0] == ("free", t), f"G2 step {t} writes the head columns while D[0] = {m.D[0]}"
            issue(chk, lambda: None)
            if j == 3:
                m.pipe.append(lambda t=t: setattr(m, "head", ("G2", t)))
                m.pipe.append(m.head_ready.arrive)
            yield


def ro_epilogue(m, half, steps):
    who = f"epi{half}"
    for t in range(steps):
        for layer, (gemm, dst, kind, prev_kind) in enumerate((("G0", m.A1, "h1", "h1"), ("G1", m.A2, "h2", "h2"))):
            for cc in range(PARTS):
                c = cc * 2 + half
                yield from m.wait(who, m.d_ready[cc], layer)
                assert m.D[cc] == (gemm, t), f"E{layer} step {t} reads D[{cc}] = {m.D[cc]}, wants {gemm}"
                yield
                assert dst[c] is None or dst[c] == (prev_kind, t - 1), f"E{layer} step {t} overwrites {kind} chunk {c} = {dst[c]}"
                dst[c] = (kind, t)
                if m.chunk_done[cc].pending == 1:
                    m.D[cc] = ("free", t)
                m.chunk_done[cc].arrive()
                yield
        if half == 0:
            yield from m.wait(who, m.head_ready, t & 1)
            assert m.head == ("G2", t), f"E2 step {t} reads head = {m.head}"
            m.head = ("read", t)
            m.head_done.arrive()
            yield


def run_forward(seed, steps):
    rng = random.Random(seed)
    m = Model(steps, rng)
    actors = {"prod": ro_producer(m, steps), "mma": ro_mma(m, steps), "epi0": ro_epilogue(m, 0, steps), "epi1": ro_epilogue(m, 1, steps)}
    idle = 0
    bars = m.d_ready + m.chunk_done + [m.head_ready, m.head_done] + m.x_full + m.x_empty
    while actors or m.pipe:
        pick = rng.choice(list(actors) + (["pipe"] if m.pipe else []))
        before = (tuple(b.phase for b in bars), len(m.pipe))
        if pick == "pipe":
            m.pipe_step()
        else:
            try:
                next(actors[pick])
            except StopIteration:
                del actors[pick]
        idle = idle + 1 if before == (tuple(b.phase for b in bars), len(m.pipe)) else 0
        assert idle < 20000, f"deadlock (seed {seed}): still running {sorted(actors)}, pipe {len(m.pipe)}"
    assert m.head == ("read", steps - 1)


@pytest.mark.parametrize("steps", [1, 2, 5])
def test_forward_and_rollout_barrier_protocol_under_random_schedules(steps):
    for seed in range(150):
        run_forward(seed * 6151 + steps, steps)
