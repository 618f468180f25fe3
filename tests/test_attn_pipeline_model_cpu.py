"""CPU model of the producer / MMA / element-wise pipelines of the tcgen05 attention kernels (attn_fwd_tc5.cu, attn_bwd_tc5.cu).

The kernels are three kinds of agents that only meet at mbarriers: a TMA producer, one MMA-issuing thread (whose MMAs and commits retire
IN ORDER) and element-wise warpgroups.  This model replays exactly the wait / arrive / commit sequence each agent executes in the CUDA
source -- same barrier arrays, same ring depths, same parities -- under randomly interleaved scheduling, and checks what the hardware
would not tell us politely:
  * no deadlock for any tile count (including 0, 1, 2: the drain iterations of the software-pipelined MMA loops),
  * a ring stage / TMEM buffer is never overwritten before its last reader has retired (the in-order MMA pipe is modelled explicitly),
  * every tile is consumed exactly once, in order.
It is a model of the protocol, not of the arithmetic (the GPU tests cover that)."""
import random

import pytest


class Bar:
    """mbarrier with `count` arrivals per phase; wait(parity) passes once the phase with that parity has completed."""
    def __init__(self, count=1):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        if self.pending == 0:
            self.pending, self.phase = self.count, self.phase + 1

    def done(self, parity):                 # try_wait.parity semantics: the CURRENT phase's parity differs from the waited one
        return (self.phase & 1) != parity


def run(agents, seed, max_steps=200000):
    """agents: generators yielding ('wait', bar, parity) | ('mma', fn) | None.  MMA-pipe work (`fn`) retires in issue order, lazily."""
    rng = random.Random(seed)
    live = {i: a for i, a in enumerate(agents)}
    blocked = {}
    pipe = []                                # in-order tensor pipe: queued closures (MMAs and commits)
    for _ in range(max_steps):
        if not live:
            while pipe:
                pipe.pop(0)()
            return True
        # retire some queued tensor-pipe work (in order)
        for _ in range(rng.randint(0, 3)):
            if pipe:
                pipe.pop(0)()
        i = rng.choice(list(live))
        if i in blocked:
            bar, parity = blocked[i]
            if not bar.done(parity):
                if all(j in blocked and not blocked[j][0].done(blocked[j][1]) for j in live) and not pipe:
                    raise AssertionError(f"deadlock: agents {sorted(live)} all blocked")
                continue
            del blocked[i]
        try:
            ev = next(live[i])
        except StopIteration:
            del live[i]
            continue
        if ev is None:
            continue
        if ev[0] == "wait":
            blocked[i] = (ev[1], ev[2])
        elif ev[0] == "mma":
            pipe.append(ev[1])
    raise AssertionError("did not finish")


class Res:
    """A ring stage or TMEM buffer: tracks which tile it holds and whether that content has been fully consumed."""
    def __init__(self, name):
        self.name, self.tile, self.readers_left = name, None, 0

    def write(self, tile, readers):
        assert self.readers_left == 0, f"{self.name}: tile {tile} overwrites tile {self.tile} with {self.readers_left} reader(s) outstanding"
        self.tile, self.readers_left = tile, readers

    def read(self, tile):
        assert self.tile == tile, f"{self.name}: expected tile {tile}, holds {self.tile}"
        assert self.readers_left > 0
        self.readers_left -= 1


# ------------------------------------------------------------------------------------------------------------- dq kernel
def model_dq(n_tiles, seed, NSTK=4, NSTV=3, NBUF=3):
    """attn_bwd_dq_kernel: K ring NSTK, V ring NSTV, NBUF score buffers, two element-wise groups alternating tiles."""
    k_full, k_empty = [Bar() for _ in range(NSTK)], [Bar() for _ in range(NSTK)]
    v_full, v_empty = [Bar() for _ in range(NSTV)], [Bar() for _ in range(NSTV)]
    sdp_full, ds_full = [Bar() for _ in range(NBUF)], [Bar(4) for _ in range(NBUF)]
    dq_final = Bar()
    Ks, Vs = [Res(f"K{s}") for s in range(NSTK)], [Res(f"V{s}") for s in range(NSTV)]
    S = [Res(f"S/dP{b}") for b in range(NBUF)]            # scores, then dS in place
    consumed = []

    def producer():
        sk = sv = 0; phk = phv = 0
        for t in range(n_tiles):
            yield ("wait", k_empty[sk], phk ^ 1)
            Ks[sk].write(t, 2); k_full[sk].arrive()        # readers: S = Q K^T and dQ += dS K
            yield ("wait", v_empty[sv], phv ^ 1)
            Vs[sv].write(t, 1); v_full[sv].arrive()
            sk += 1; sv += 1
            if sk == NSTK: sk, phk = 0, phk ^ 1
            if sv == NSTV: sv, phv = 0, phv ^ 1
            yield None

    def mma():
        sk = sv = su = 0; phk = phv = 0
        for t in range(n_tiles + 2):
            if t < n_tiles:
                b3 = t % NBUF
                yield ("wait", k_full[sk], phk)
                yield ("wait", v_full[sv], phv)
                def f(t=t, b3=b3, sk=sk, sv=sv):
                    Ks[sk].read(t); Vs[sv].read(t); S[b3].write(t, 5)        # 4 element-wise warps + the dQ MMA read it
                    sdp_full[b3].arrive(); v_empty[sv].arrive()
                yield ("mma", f)
                sk += 1; sv += 1
                if sk == NSTK: sk, phk = 0, phk ^ 1
                if sv == NSTV: sv, phv = 0, phv ^ 1
            if t >= 2:
                u = t - 2; b3 = u % NBUF
                yield ("wait", ds_full[b3], (u // NBUF) & 1)
                def g(u=u, b3=b3, su=su):
                    S[b3].read(u); Ks[su].read(u); consumed.append(u); k_empty[su].arrive()
                yield ("mma", g)
                su = (su + 1) % NSTK
        yield ("mma", dq_final.arrive)

    def ew(g, w):
        for t in range(g, n_tiles, 2):
            b3 = t % NBUF
            yield ("wait", sdp_full[b3], (t // NBUF) & 1)
            S[b3].read(t)
            yield None
            ds_full[b3].arrive()
        if n_tiles > 0:
            yield ("wait", dq_final, 0)

    agents = [producer(), mma()] + [ew(g, w) for g in range(2) for w in range(4)]
    assert run(agents, seed)
    assert consumed == list(range(n_tiles))


# ------------------------------------------------------------------------------------------------------------ dk/dv kernel
def model_dkv(iters, seed, NST=3):
    """attn_bwd_dkv_kernel: Q / dO ring NST, two score buffers, two element-wise groups alternating iterations."""
    q_full, do_full, qdo_empty = [Bar() for _ in range(NST)], [Bar() for _ in range(NST)], [Bar() for _ in range(NST)]
    sdp_full, pds_full = [Bar() for _ in range(2)], [Bar(4) for _ in range(2)]
    acc_final = Bar()
    QD = [Res(f"Q/dO{s}") for s in range(NST)]
    S = [Res(f"S^T/dP^T{b}") for b in range(2)]
    consumed = []

    def producer():
        s = ph = 0
        for it in range(iters):
            yield ("wait", qdo_empty[s], ph ^ 1)
            QD[s].write(it, 2); q_full[s].arrive(); do_full[s].arrive()
            s += 1
            if s == NST: s, ph = 0, ph ^ 1
            yield None

    def mma():
        s = ph = su = 0
        for it in range(iters + 1):
            if it < iters:
                yield ("wait", q_full[s], ph)
                yield ("wait", do_full[s], ph)
                def f(it=it, s=s):
                    QD[s].read(it); S[it & 1].write(it, 5); sdp_full[it & 1].arrive()
                yield ("mma", f)
                s += 1
                if s == NST: s, ph = 0, ph ^ 1
            if it >= 1:
                u = it - 1
                yield ("wait", pds_full[u & 1], (u >> 1) & 1)
                def g(u=u, su=su):
                    S[u & 1].read(u); QD[su].read(u); consumed.append(u); qdo_empty[su].arrive()
                yield ("mma", g)
                su = (su + 1) % NST
        yield ("mma", acc_final.arrive)

    def ew(g, w):
        for it in range(g, iters, 2):
            yield ("wait", sdp_full[g], (it >> 1) & 1)
            S[g].read(it)
            yield None
            pds_full[g].arrive()
        if iters > 0:
            yield ("wait", acc_final, 0)

    assert run([producer(), mma()] + [ew(g, w) for g in range(2) for w in range(4)], seed)
    assert consumed == list(range(iters))


# --------------------------------------------------------------------------------------------------------------- forward
def model_fwd(n_tiles, seed, NST=2):
    """attn_fwd_tc5_kernel: K / V rings NST, two score buffers, one softmax group (4 warps); QK_t is issued before PV_{t-1}."""
    k_full, k_empty = [Bar() for _ in range(NST)], [Bar() for _ in range(NST)]
    v_full, v_empty = [Bar() for _ in range(NST)], [Bar() for _ in range(NST)]
    s_full, p_full, pv_done = [Bar() for _ in range(2)], [Bar(4) for _ in range(2)], [Bar() for _ in range(2)]
    Ks, Vs = [Res(f"K{s}") for s in range(NST)], [Res(f"V{s}") for s in range(NST)]
    S = [Res(f"S/P{b}") for b in range(2)]
    consumed = []

    def producer():
        s = ph = 0
        for t in range(n_tiles):
            yield ("wait", k_empty[s], ph ^ 1)
            Ks[s].write(t, 1); k_full[s].arrive()
            yield ("wait", v_empty[s], ph ^ 1)
            Vs[s].write(t, 1); v_full[s].arrive()
            s += 1
            if s == NST: s, ph = 0, ph ^ 1

    def mma():
        s = ph = sp = php = 0
        for t in range(n_tiles + 1):
            if t < n_tiles:
                yield ("wait", k_full[s], ph)
                def f(t=t, s=s):
                    Ks[s].read(t); S[t & 1].write(t, 5); s_full[t & 1].arrive(); k_empty[s].arrive()
                yield ("mma", f)
                s += 1
                if s == NST: s, ph = 0, ph ^ 1
            if t >= 1:
                u = t - 1
                yield ("wait", p_full[u & 1], (u >> 1) & 1)
                yield ("wait", v_full[sp], php)
                def g(u=u, sp=sp):
                    S[u & 1].read(u); Vs[sp].read(u); consumed.append(u); v_empty[sp].arrive(); pv_done[u & 1].arrive()
                yield ("mma", g)
                sp += 1
                if sp == NST: sp, php = 0, php ^ 1

    def softmax(w, rescale_at):
        for t in range(n_tiles):
            yield ("wait", s_full[t & 1], (t >> 1) & 1)
            S[t & 1].read(t)
            if t > 0 and t in rescale_at:                    # lazy rescale of O: needs PV_{t-1} retired
                yield ("wait", pv_done[(t - 1) & 1], ((t - 1) >> 1) & 1)
            p_full[t & 1].arrive()
        if n_tiles > 0:
            yield ("wait", pv_done[(n_tiles - 1) & 1], ((n_tiles - 1) >> 1) & 1)

    rng = random.Random(seed)
    rescale = {t for t in range(n_tiles) if rng.random() < 0.3}
    assert run([producer(), mma()] + [softmax(w, rescale) for w in range(4)], seed)
    assert consumed == list(range(n_tiles))


@pytest.mark.parametrize("n", [0, 1, 2, 3, 4, 5, 7, 12, 37])
def test_attention_pipelines_are_deadlock_and_hazard_free(n):
    for seed in range(6):
        model_dq(n, seed)
        model_dkv(n, seed)
        model_fwd(n, seed)


def test_model_detects_a_too_shallow_k_ring():
    """The dq kernel holds a K tile until dQ += dS K of ITS tile retires, two tiles after its scores were issued: with three score
    buffers a 2-stage K ring must deadlock or be flagged (this is why the kernel uses 4 stages) -- the model has to notice."""
    with pytest.raises(AssertionError):
        for seed in range(8):
            model_dq(9, seed, NSTK=2)
