"""Randomised interleaving check of the mbarrier protocol of csrc/apply_tc3.cu — THE DEFAULT APPLY KERNEL — as committed at the end of
round 1 (transform loop that waits for tcgen05.st inside the iteration, two MMA warps, early addend prefetch into released raw stages,
Qt slot 1 outside the E ring).  CPU model, no GPU; same construction as scripts/protocol_sim.py:

  * every warp role is a coroutine following the kernel's loops and barrier waits line by line (8 transform / conversion / epilogue
    warps in two groups, W-TMA warp, E/Qt-TMA warp, two MMA warps);
  * TMA loads complete at random later times in any order, TMA stores finish READING shared memory at random later times in issue
    order (cp.async.bulk.wait_group.read semantics), tensor-core work executes asynchronously in issue order and its commits arrive when
    it has executed;
  * shared memory and tensor memory are modelled with the kernel's ALIASING: box pair b < 5 = raw stage b, box pair 5 = E stages 0-1,
    Qt slot 0 = E stage 2 (R = 64), P_lo = A stage 0, accumulators = A stages 1-2.

Invariants checked at every access: a reader finds the content it expects (chunk / unit id and kind), a writer never overwrites content
that still has pending readers, an MMA finds its A rows from all four warps of its group and its B tile, nothing deadlocks.
      python scripts/protocol_sim_tc3.py [runs]
"""
import random
import sys

NRAW, NSA, NE, NB, NACC, NQ, PW, NBLK = 5, 3, 3, 6, 4, 2, 8, 2


class Barrier:
    def __init__(self, count, name=""):
        self.count, self.pending, self.phase, self.name = count, count, 0, name

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, f"{self.name}: more arrivals than expected in one phase"
        if self.pending == 0:
            self.pending, self.phase = self.count, self.phase ^ 1

    def passed(self, parity):
        return self.phase != parity


class Region:
    """A piece of shared or tensor memory: what it holds and who still has to read that."""
    def __init__(self, name):
        self.name, self.content, self.readers_left = name, None, 0

    def write(self, content, readers, who):
        assert self.readers_left == 0, f"{who} overwrote {self.name} = {self.content} with {content}: {self.readers_left} reader(s) still pending"
        self.content, self.readers_left = content, readers

    def read(self, expect, who):
        assert self.content == expect, f"{who} read {self.name} = {self.content}, expected {expect}"
        assert self.readers_left > 0, f"{who} read {self.name} = {self.content} more often than planned"
        self.readers_left -= 1


class Sim:
    def __init__(self, n_chunks, n_act, rng, pipelined=False):
        self.n, self.n_act, self.rng, self.pipelined = n_chunks, n_act, rng, pipelined
        B = Barrier
        self.raw_full = [B(1 + n_act, f"raw_full{r}") for r in range(NRAW)]
        self.raw_empty = [B(PW, f"raw_empty{r}") for r in range(NRAW)]
        self.a_full = [B(PW, f"a_full{s}") for s in range(NSA)]
        self.a_empty = [B(NBLK, f"a_empty{s}") for s in range(NSA)]
        self.e_full = [B(1 + 2, f"e_full{s}") for s in range(NE)]
        self.e_empty = [B(NBLK, f"e_empty{s}") for s in range(NE)]
        self.p_full, self.p_ready = B(NBLK, "p_full"), B(PW, "p_ready")
        self.q_full = [B(1 + 1, f"q_full{t}") for t in range(NQ)]            # all tiles of a unit modelled as one load
        self.q_empty = [B(NBLK, f"q_empty{t}") for t in range(NQ)]
        self.acc_full = [[B(1, f"acc_full{a}{g}") for g in range(NBLK)] for a in range(NACC)]
        self.acc_empty = [[B(4, f"acc_empty{a}{g}") for g in range(NBLK)] for a in range(NACC)]
        self.box_full = [B(1 + n_act, f"box_full{b}") for b in range(NB)]
        self.box_ready = [B(PW, f"box_ready{b}") for b in range(NB)]
        # shared memory with the kernel's aliasing (R = 64)
        self.slot = [[Region(f"smem pair {i} block {g}") for g in range(NBLK)] for i in range(NB)]      # raw stage i / box pair i (i < 5); box pair 5 below
        self.E = [Region(f"E stage {s}") for s in range(NE)]
        self.Q1 = Region("Qt slot 1")
        # tensor memory: A stages per (stage, group, warp); P per group; accumulators per (a, g)
        self.A = [[[Region(f"A stage {s} group {g} warp {w}") for w in range(4)] for g in range(NBLK)] for s in range(NSA)]
        self.P = [Region(f"P group {g}") for g in range(NBLK)]
        self.ACC = [[Region(f"acc {a} group {g}") for g in range(NBLK)] for a in range(NACC)]
        self.tma_loads, self.store_groups, self.tensor_queue = [], [], []
        self.phase_a_done = [0, 0]
        self.finished_units = 0

    # ---- aliasing helpers ----------------------------------------------------------------------------------------------
    def box_regions(self, b, g):
        """Regions a box of pair b, block g covers."""
        if b < NRAW:
            return [self.slot[b][g]]
        return [self.slot[5][g], self.E[g]]                       # pair 5 = [160K, 192K) = E stage 0 (block 0) and E stage 1 (block 1)

    def qt_region(self, t):
        return self.E[2] if t == 0 else self.Q1

    # ---- engines ---------------------------------------------------------------------------------------------------------
    def tma_engine(self):
        while True:
            if self.tma_loads:
                fn = self.tma_loads.pop(self.rng.randrange(len(self.tma_loads)))
                fn()
            yield

    def store_engine(self):
        while True:
            if self.store_groups and self.rng.random() < 0.5:
                grp = self.store_groups[0]
                if not grp["done"]:
                    for reg, expect in grp["reads"]:
                        reg.read(expect, f"TMA store of unit {grp['unit']}")
                    grp["done"] = True
                self.store_groups.pop(0)
                self.finished_units += 1
            yield

    def tensor_engine(self):
        while True:
            if self.tensor_queue and self.rng.random() < 0.7:
                fn = self.tensor_queue.pop(0)
                fn()
            yield

    def wait(self, bar, parity):
        while not bar.passed(parity):
            yield

    # ---- warp roles ---------------------------------------------------------------------------------------------------
    def transform(self, g, w):
        who = f"transform warp g{g}w{w}"
        live = g < self.n_act
        pending = None                                          # pipelined shape: A stage stored but not yet published
        for c in range(self.n):
            r, s = c % NRAW, c % NSA
            yield from self.wait(self.raw_full[r], (c // NRAW) & 1)
            if live:
                self.slot[r][g].read(("raw", c), who)
            yield
            if self.pipelined:                                      # the shape the round's last GPU runs validated and profiled
                self.raw_empty[r].arrive()
                if pending is not None:
                    self.a_full[pending].arrive(); pending = None
                yield
            yield from self.wait(self.a_empty[s], ((c // NSA) & 1) ^ 1)
            self.A[s][g][w].write(("A", c), 1 if live else 0, who)
            yield
            if self.pipelined:
                pending = s
            else:
                self.a_full[s].arrive(); self.raw_empty[r].arrive()
                yield
        if pending is not None:
            self.a_full[pending].arrive()
            yield
        yield from self.wait(self.p_full, 0)
        if live and w == 0:
            self.P[g].read(("P",), who)                             # P accumulator final -> split in place (one logical reader per group)
            self.P[g].write(("Psplit",), self.n, who)               # read by every phase-B unit of this group
        # P_lo goes into A stage 0 of the group: its last phase-A readers must be done (p_full guarantees it)
        if live:
            self.A[0][g][w].write(("Plo",), 0, who)
        yield
        self.p_ready.arrive()
        yield
        for u in range(self.n):
            b, a = u % NB, u % NACC
            yield from self.wait(self.acc_full[a][g], (u // NACC) & 1)
            if live and w == 0:
                self.ACC[a][g].read(("acc", u), who)
            yield
            if True:
                self.acc_empty[a][g].arrive()
            yield from self.wait(self.box_full[b], (u // NB) & 1)
            if live and w == 0:                                     # the group's four warps update disjoint rows of the same box: one logical access
                for reg in self.box_regions(b, g):
                    reg.read(("add", u), who)
                    reg.write(("new", u), 1, who)                   # read once more: by the TMA store
            yield
            self.box_ready[b].arrive()
            yield

    def w_tma(self):
        who = "W-TMA warp"
        n_act = self.n_act

        def load_raw(r, c):
            self.raw_full[r].arrive()
            for g in range(n_act):
                def done(g=g):
                    self.slot[r][g].write(("raw", c), 4, "TMA load of raw chunk %d" % c)
                    self.raw_full[r].arrive()
                self.tma_loads.append(done)

        def load_box(unit, bx):
            self.box_full[bx].arrive()
            for g in range(n_act):
                def done(g=g):
                    for reg in self.box_regions(bx, g):
                        reg.write(("add", unit), 1, "TMA load of the addend of unit %d" % unit)
                    self.box_full[bx].arrive()
                self.tma_loads.append(done)

        for c in range(self.n):
            r = c % NRAW
            yield from self.wait(self.raw_empty[r], ((c // NRAW) & 1) ^ 1)
            load_raw(r, c)
            yield
        for u in range(min(NB, self.n)):
            if u < NRAW:
                uses = (self.n - u + NRAW - 1) // NRAW
                yield from self.wait(self.raw_empty[u], (uses - 1) & 1)
            else:
                yield from self.wait(self.p_full, 0)
            load_box(u, u)
            yield
        issued = []
        for u in range(self.n):
            b = u % NB
            yield from self.wait(self.box_ready[b], (u // NB) & 1)
            nu = u - 1 + NB
            reload = u >= 1 and nu < self.n
            grp = {"unit": u, "done": False, "reads": [(reg, ("new", u)) for g in range(n_act) for reg in self.box_regions(b, g)]}
            self.store_groups.append(grp); issued.append(grp)
            yield
            if reload:                                              # wait_group.read 1: every store but the newest has read its boxes
                while any(not gr["done"] for gr in issued[:-1]):
                    yield
                load_box(nu, nu % NB)
                yield
        while any(not gr["done"] for gr in issued):                 # wait_group.read 0
            yield

    def e_tma(self):
        for c in range(self.n):
            s = c % NE
            yield from self.wait(self.e_empty[s], ((c // NE) & 1) ^ 1)
            self.e_full[s].arrive()
            for part in range(2):
                def done(s=s, c=c, part=part):
                    if part == 0:
                        self.E[s].write(("E", c), self.n_act, "TMA load of E chunk %d" % c)
                    self.e_full[s].arrive()
                self.tma_loads.append(done)
            yield
        for u in range(self.n):
            t = (u + 1) % NQ
            if u == 1:
                yield from self.wait(self.p_full, 0)
            yield from self.wait(self.q_empty[t], ((u // NQ) & 1) ^ 1)
            self.q_full[t].arrive()

            def done(t=t, u=u):
                self.qt_region(t).write(("Q", u), self.n_act, "TMA load of Qt unit %d" % u)
                self.q_full[t].arrive()
            self.tma_loads.append(done)
            yield

    def mma(self, g):
        who = f"MMA warp {g}"
        act = g < self.n_act
        for c in range(self.n):
            s, se = c % NSA, c % NE
            yield from self.wait(self.e_full[se], (c // NE) & 1)
            yield from self.wait(self.a_full[s], (c // NSA) & 1)
            last = c == self.n - 1
            if act:
                def run(c=c, s=s, se=se, last=last):
                    for w in range(4):
                        self.A[s][g][w].read(("A", c), f"phase-A MMA g{g} chunk {c}")
                    self.E[se].read(("E", c), f"phase-A MMA g{g} chunk {c}")
                    if last:
                        self.P[g].write(("P",), 1, who)
                    self.a_empty[s].arrive(); self.e_empty[se].arrive()
                    if last:
                        self.p_full.arrive()
                self.tensor_queue.append(run)
            else:
                self.a_empty[s].arrive(); self.e_empty[se].arrive()
                if last:
                    self.p_full.arrive()
            yield
        yield from self.wait(self.p_ready, 0)
        for u in range(self.n):
            a, t = u % NACC, (u + 1) % NQ
            yield from self.wait(self.acc_empty[a][g], ((u // NACC) & 1) ^ 1)
            yield from self.wait(self.q_full[t], (u // NQ) & 1)
            if act:
                def run(u=u, a=a, t=t):
                    self.P[g].read(("Psplit",), f"phase-B MMA g{g} unit {u}")
                    self.qt_region(t).read(("Q", u), f"phase-B MMA g{g} unit {u}")
                    # accumulators live in A stages 1-2 (a = 0,1 -> stage 1; a = 2,3 -> stage 2): nothing of phase A may be pending there
                    st = 1 + a // 2
                    for gg in range(NBLK):
                        for w in range(4):
                            assert self.A[st][gg][w].readers_left == 0, f"accumulator {a} of group {g} written over A stage {st} with pending readers"
                    self.ACC[a][g].write(("acc", u), 1, who)
                    self.q_empty[t].arrive(); self.acc_full[a][g].arrive()
                self.tensor_queue.append(run)
            else:
                self.q_empty[t].arrive(); self.acc_full[a][g].arrive()
            yield

    def run(self):
        agents = {"w_tma": self.w_tma(), "e_tma": self.e_tma(), "mma0": self.mma(0), "mma1": self.mma(1),
                  "tma_engine": self.tma_engine(), "store_engine": self.store_engine(), "tensor_engine": self.tensor_engine()}
        for g in range(NBLK):
            for w in range(4):
                agents[f"t{g}{w}"] = self.transform(g, w)
        finite = {k for k in agents if not k.endswith("engine")}
        steps = 0
        # skewed timing: every agent gets a random speed for the whole run (log-uniform over three decades), so schedules in which one
        # warp, one engine or one role is far slower than the rest are as common as balanced ones
        speed = {k: 10.0 ** self.rng.uniform(-3, 0) for k in agents}
        while finite:
            names = list(agents)
            name = self.rng.choices(names, weights=[speed[k] for k in names])[0]
            try:
                next(agents[name])
            except StopIteration:
                del agents[name]; finite.discard(name)
            steps += 1
            assert steps < 400_000_000, f"no termination (deadlock?) with {sorted(finite)} still running"
        assert self.finished_units == self.n, (self.finished_units, self.n)


def main():
    runs = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = random.Random(0)
    for pipelined in (False, True):          # False: the committed transform loop; True: the software-pipelined one the last GPU runs used
        for n_act in (2, 1):
            for n_chunks in (1, 2, 3, 4, 5, 6, 7, 8, 11, 12, 16, 24):
                for _ in range(runs):
                    Sim(n_chunks, n_act, rng, pipelined).run()
            print(f"apply_tc3 protocol ({'pipelined' if pipelined else 'committed'} transform loop), {n_act} active row block(s): "
                  f"{runs} random schedules x 12 chunk counts (K = 32 .. 768): no violation")


if __name__ == "__main__":
    main()
