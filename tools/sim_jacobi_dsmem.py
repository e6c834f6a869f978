"""CPU model of the hand-over protocol of jacobi_cluster_kernel (csrc/tail.cu): slots, ready / consumed mbarrier phases,
global-memory path across cluster boundaries, random interleaving of the CTAs.  Checks: no deadlock; every pull / load
delivers the block the permutation says sits at that position; no slot is overwritten before its consumer has read it;
every pair of blocks meets exactly once per sweep; at exit the home copy of every block is its latest version.
    python tools/sim_jacobi_dsmem.py [NP] [cluster] [sweeps] [seeds]
"""
import random
import sys


class MBar:
    """mbarrier with arrival count 1: a phase completes on every arrive; wait(parity) succeeds when the phase with that
    parity has completed, i.e. when the current phase parity differs (only the last phase is visible)."""

    def __init__(self):
        self.phase = 0

    def arrive(self):
        self.phase += 1

    def test(self, parity):
        return (self.phase & 1) != parity


class CTA:
    def __init__(self, sim, cta):
        self.sim, self.cta = sim, cta
        self.slots = [None, None, None]          # block content (block id) physically in each slot
        self.sblk = [-1, -1, -1]
        self.ready = [MBar(), MBar()]
        self.cons = [MBar(), MBar()]
        self.reading = {}                        # slot -> consumer currently reading it (overwrite check)

    def run(self):
        sim, cta = self.sim, self.cta
        NP, N, CS = sim.NP, sim.N, sim.CS
        active = cta < NP
        has = [active and cta > 0, active and cta + 1 < NP]
        dsm = [has[0] and (cta - 1) // CS == cta // CS, has[1] and (cta + 1) // CS == cta // CS]
        nb = [cta - 1, cta + 1]
        perm = list(range(N))
        nready, ncons = [0, 0], [0, 0]
        gone = -1
        last = active and cta == NP - 1
        gstep = 0
        for sweep in range(sim.sweeps):
            for st in range(N):
                t = gstep
                if active and (t % 2 == 0):
                    pL, pR = 2 * cta, 2 * cta + 1
                elif active and cta < NP - 1:
                    pL, pR = 2 * cta + 1, 2 * cta + 2
                else:
                    pL = pR = -1
                if gone >= 0:
                    self.sblk[gone] = -1
                    gone = -1
                if pL >= 0:
                    s_new = (t + 1) % 3 if last else t % 3
                    s_old = (t + 2) % 3
                    pos_new, pos_old = (pR, pL) if t & 1 else (pL, pR)
                    if t == 0:
                        for s, pos in ((s_old, pos_old), (s_new, pos_new)):
                            self.land(s, sim.home[perm[pos]])
                            self.sblk[s] = perm[pos]
                    else:
                        dprev = t & 1
                        if t >= 2 and dsm[dprev]:
                            while not self.cons[dprev].test(ncons[dprev] & 1):
                                yield "cons"
                            ncons[dprev] += 1
                        din = 1 if t & 1 else 0
                        if has[din]:
                            if dsm[din]:
                                while not self.ready[din].test(nready[din] & 1):
                                    yield "ready"
                                nready[din] += 1
                                prod = sim.ctas[nb[din]]
                                src = (t + 1) % 3
                                prod.reading[src] = cta
                                yield "pull"                                   # the copy takes time
                                self.land(s_new, prod.slots[src])
                                del prod.reading[src]
                                prod.cons[1 - din].arrive()
                            else:
                                while sim.bstep[perm[pos_new]] < t:
                                    yield "flag"
                                yield "load"
                                self.land(s_new, sim.home[perm[pos_new]])
                            self.sblk[s_new] = perm[pos_new]
                    a, b = self.slots[s_new], self.slots[s_old]
                    # the right blocks, in their latest versions (a stale copy would carry an older version)
                    assert a == ("blk", perm[pos_new], sim.latest[perm[pos_new]]), (cta, t, a, perm[pos_new])
                    assert b == ("blk", perm[pos_old], sim.latest[perm[pos_old]]), (cta, t, b, perm[pos_old])
                    assert self.sblk[s_new] == perm[pos_new] and self.sblk[s_old] == perm[pos_old]
                    sim.met(sweep, perm[pos_new], perm[pos_old])
                    for s_ in (s_new, s_old):                                  # the rotations rewrite both blocks
                        bid = self.slots[s_][1]
                        sim.latest[bid] += 1
                        self.slots[s_] = ("blk", bid, sim.latest[bid])
                    assert s_new not in self.reading and s_old not in self.reading, ("rotating a slot that is being read", cta, t)
                    yield "rounds"
                    assert s_new not in self.reading and s_old not in self.reading, ("rotating a slot that is being read", cta, t)
                npairs, o = (NP, 0) if gstep % 2 == 0 else (NP - 1, 1)
                for i in range(npairs):
                    perm[2 * i + o], perm[2 * i + o + 1] = perm[2 * i + o + 1], perm[2 * i + o]
                if pL >= 0:
                    dout = 1 if t & 1 else 0
                    s_old = (t + 2) % 3
                    if has[dout]:
                        if dsm[dout]:
                            sim.ctas[nb[dout]].ready[1 - dout].arrive()
                            gone = s_old
                        else:
                            blk = self.sblk[s_old]
                            yield "store"
                            sim.home[blk] = self.slots[s_old]
                            sim.bstep[blk] = t + 1
                            self.sblk[s_old] = -1
                gstep += 1
            while not sim.grid_barrier(cta, sweep):
                yield "grid"
        for s in range(3):
            if self.sblk[s] >= 0:
                sim.final_store(self.sblk[s], self.slots[s])

    def land(self, slot, content):
        assert slot not in self.reading, ("overwrite while being read", self.cta, slot)
        assert content is not None
        self.slots[slot] = content


class Sim:
    def __init__(self, NP, CS, sweeps, seed):
        self.NP, self.N, self.CS, self.sweeps = NP, 2 * NP, CS, sweeps
        self.G = -(-NP // CS) * CS
        self.rng = random.Random(seed)
        self.home = {b: ("blk", b, 0) for b in range(self.N)}
        self.latest = {b: 0 for b in range(self.N)}
        self.bstep = {b: 0 for b in range(self.N)}
        self.pairs = [set() for _ in range(sweeps)]
        self.arrived = [set() for _ in range(sweeps)]
        self.stored = {}
        self.ctas = [CTA(self, c) for c in range(self.G)]

    def met(self, sweep, a, b):
        key = (min(a, b), max(a, b))
        assert key not in self.pairs[sweep], ("pair met twice", sweep, key)
        self.pairs[sweep].add(key)

    def grid_barrier(self, cta, sweep):
        self.arrived[sweep].add(cta)
        return len(self.arrived[sweep]) == self.G

    def final_store(self, blk, content):
        assert blk not in self.stored, ("block stored twice", blk)
        assert content[:2] == ("blk", blk)
        self.stored[blk] = content
        self.home[blk] = content

    def run(self):
        gens = {c.cta: c.run() for c in self.ctas}
        blocked = {}
        while gens:
            cta = self.rng.choice(list(gens))
            try:
                blocked[cta] = next(gens[cta])
            except StopIteration:
                del gens[cta]
                blocked.pop(cta, None)
                continue
            # deadlock detection: a long streak in which nobody makes progress is impossible here because every yield
            # either is a timing yield (always progresses next time) or re-tests a condition; bound the total work instead
            self.ticks = getattr(self, "ticks", 0) + 1
            assert self.ticks < 5_000_000, ("no progress (deadlock?)", blocked)
        for sw in range(self.sweeps):
            assert len(self.pairs[sw]) == self.NP * (self.N - 1), (sw, len(self.pairs[sw]), self.NP * (self.N - 1))
        # every block's home copy is its latest version (stored at exit, or handed over through global memory last)
        for b in range(self.N):
            assert self.home[b] == ("blk", b, self.latest[b]), (b, self.home[b], self.latest[b])


def main():
    NP = int(sys.argv[1]) if len(sys.argv) > 1 else 7
    CS = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    sweeps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    seeds = int(sys.argv[4]) if len(sys.argv) > 4 else 20
    for seed in range(seeds):
        Sim(NP, CS, sweeps, seed).run()
    print(f"ok: NP={NP} cluster={CS} sweeps={sweeps} seeds={seeds}: every pair met once per sweep, every block home in its last version, "
          "no deadlock, no early overwrite")


if __name__ == "__main__":
    main()
