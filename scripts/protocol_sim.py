"""Randomised interleaving check of the mbarrier protocol of csrc/apply_gemm3x.cu (rings, phases, who arrives where) — a CPU model, no
GPU.  Every warp role of the kernel is a coroutine that follows the kernel's loop line by line; TMA loads and tensor-core MMAs are
asynchronous engines that complete at random later times (MMAs in issue order, like the tensor pipe); the scheduler picks a random
runnable agent at every step.  Checked while running: a raw stage is only overwritten after all four transform warps have read it, an
A stage only after the MMAs that read it have completed, every MMA sees the A rows of all four warps and the B tile of ITS chunk, the
epilogue starts after the last MMA, and no run deadlocks.  Both loop shapes are modelled: the software-pipelined transform loop of the
first hardware runs (the one that produced the intermittent warp-0 errors, profiles/r01_gemm3x_diag_*.txt) and the current one.

A clean run says the protocol AS WRITTEN is sound under mbarrier semantics (parity waits, arrival counts); it says nothing about the
hardware rules the model does not contain (tcgen05 ordering, proxies).      python scripts/protocol_sim.py [runs]
"""
import random
import sys

NRAW, NSA, NBS, PW = 4, 4, 2, 4


class Barrier:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, "more arrivals than the barrier expects in one phase"
        if self.pending == 0:
            self.pending, self.phase = self.count, self.phase ^ 1

    def passed(self, parity):                       # mbarrier.try_wait.parity: the phase with this parity has completed
        return self.phase != parity


class Sim:
    def __init__(self, n_chunks, pipelined, rng):
        self.n, self.pipelined, self.rng = n_chunks, pipelined, rng
        self.raw_full = [Barrier(2) for _ in range(NRAW)]       # arrive.expect_tx + the bytes landing
        self.raw_empty = [Barrier(PW) for _ in range(NRAW)]
        self.a_full = [Barrier(PW) for _ in range(NSA)]
        self.a_empty = [Barrier(1) for _ in range(NSA)]
        self.b_full = [Barrier(2) for _ in range(NBS)]
        self.b_empty = [Barrier(1) for _ in range(NBS)]
        self.acc_full = Barrier(1)
        self.raw = [None] * NRAW                                 # chunk id held by the raw stage
        self.raw_readers = [set() for _ in range(NRAW)]          # warps that have read the current content
        self.A = [[None] * PW for _ in range(NSA)]               # chunk id of each warp's rows in the A stage
        self.A_busy = [0] * NSA                                  # MMAs issued on the stage and not yet completed
        self.B = [None] * NBS
        self.B_busy = [0] * NBS
        self.tma_inflight = []                                   # (kind, stage, chunk)
        self.mma_queue = []                                      # issued, not yet executed: (chunk, s, sb, commits)
        self.mma_done = 0
        self.epilogue_started = False

    # ---- asynchronous engines -------------------------------------------------------------------------------------
    def tma_engine(self):
        while True:
            if self.tma_inflight:
                kind, st, c = self.tma_inflight.pop(self.rng.randrange(len(self.tma_inflight)))     # loads complete in any order
                if kind == "raw":
                    assert len(self.raw_readers[st]) in (0, PW) or self.raw[st] is None, f"raw stage {st} overwritten while being read"
                    self.raw[st], self.raw_readers[st] = c, set()
                    self.raw_full[st].arrive()
                else:
                    assert self.B_busy[st] == 0, f"B stage {st} overwritten under a running MMA"
                    self.B[st] = c
                    self.b_full[st].arrive()
            yield

    def tensor_engine(self):
        while True:
            if self.mma_queue:
                c, s, sb, commits = self.mma_queue.pop(0)                                           # in issue order
                assert all(x == c for x in self.A[s]), f"MMA of chunk {c} read A stage {s} = {self.A[s]}"
                assert self.B[sb] == c, f"MMA of chunk {c} read B stage {sb} = {self.B[sb]}"
                self.A_busy[s] -= 1; self.B_busy[sb] -= 1
                self.mma_done += 1
                for bar in commits:
                    bar.arrive()
            yield

    # ---- warp roles (one yield per blocking point / step) -----------------------------------------------------------
    def wait(self, bar, parity):
        while not bar.passed(parity):
            yield

    def a_tma(self):
        for c in range(self.n):
            r = c % NRAW
            yield from self.wait(self.raw_empty[r], ((c // NRAW) & 1) ^ 1)
            self.raw_full[r].arrive()                            # arrive.expect_tx
            self.tma_inflight.append(("raw", r, c))
            yield

    def b_tma(self):
        for c in range(self.n):
            s = c % NBS
            yield from self.wait(self.b_empty[s], ((c // NBS) & 1) ^ 1)
            self.b_full[s].arrive()
            self.tma_inflight.append(("b", s, c))
            yield

    def transform(self, w):
        pending_store = None                                     # pipelined shape: (stage, chunk) stored but not yet published
        for c in range(self.n):
            r, s = c % NRAW, c % NSA
            yield from self.wait(self.raw_full[r], (c // NRAW) & 1)
            assert self.raw[r] == c, f"warp {w} read raw stage {r} = {self.raw[r]} for chunk {c}"
            self.raw_readers[r].add(w)
            yield
            if self.pipelined:
                self.raw_empty[r].arrive()
                if pending_store is not None:                    # wait::st of chunk c - 1, then publish it
                    self.a_full[pending_store[0]].arrive()
                    pending_store = None
                yield
            yield from self.wait(self.a_empty[s], ((c // NSA) & 1) ^ 1)
            assert self.A_busy[s] == 0, f"warp {w} overwrote A stage {s} under a running MMA (chunk {c})"
            self.A[s][w] = c
            yield
            if self.pipelined:
                pending_store = (s, c)
            else:
                self.a_full[s].arrive(); self.raw_empty[r].arrive()
                yield
        if pending_store is not None:
            self.a_full[pending_store[0]].arrive()
        yield from self.wait(self.acc_full, 0)
        assert self.mma_done == self.n, "epilogue started before the last MMA completed"
        self.epilogue_started = True

    def mma(self):
        for c in range(self.n):
            s, sb = c % NSA, c % NBS
            yield from self.wait(self.b_full[sb], (c // NBS) & 1)
            yield from self.wait(self.a_full[s], (c // NSA) & 1)
            commits = [self.a_empty[s], self.b_empty[sb]] + ([self.acc_full] if c == self.n - 1 else [])
            self.A_busy[s] += 1; self.B_busy[sb] += 1
            self.mma_queue.append((c, s, sb, commits))
            yield

    def run(self):
        agents = {"a_tma": self.a_tma(), "b_tma": self.b_tma(), "mma": self.mma(), "tma_engine": self.tma_engine(), "tensor_engine": self.tensor_engine()}
        for w in range(PW):
            agents[f"t{w}"] = self.transform(w)
        finite = {k for k in agents if not k.endswith("engine")}
        idle = 0
        # skewed timing: every agent gets a random speed for the whole run (log-uniform over three decades)
        speed = {k: 10.0 ** self.rng.uniform(-3, 0) for k in agents}
        while finite:
            names = list(agents)
            name = self.rng.choices(names, weights=[speed[k] for k in names])[0]
            before = (self.mma_done, len(self.tma_inflight), len(self.mma_queue))
            try:
                next(agents[name])
            except StopIteration:
                del agents[name]; finite.discard(name)
            idle = idle + 1 if before == (self.mma_done, len(self.tma_inflight), len(self.mma_queue)) else 0
            assert idle < 20000000, f"no progress (deadlock?) with {sorted(finite)} still running"
        assert self.epilogue_started


class SimSS:
    """csrc/apply_gemm3x_ss.cu: 3 stages of {raw = hi | lo} in shared memory, released by the MMA warp's commit."""
    NST = 3

    def __init__(self, n_chunks, rng):
        self.n, self.rng = n_chunks, rng
        N = self.NST
        self.raw_full = [Barrier(2) for _ in range(N)]
        self.lo_full = [Barrier(PW) for _ in range(N)]
        self.st_empty = [Barrier(1) for _ in range(N)]
        self.b_full = [Barrier(2) for _ in range(NBS)]
        self.b_empty = [Barrier(1) for _ in range(NBS)]
        self.acc_full = Barrier(1)
        self.raw = [None] * N
        self.lo = [[None] * PW for _ in range(N)]
        self.busy = [0] * N                                       # MMAs issued on the stage (hi and lo tiles) and not yet completed
        self.lo_pending = [0] * N                                 # lo warps that still have to read the raw tile / write their lo rows
        self.B, self.B_busy = [None] * NBS, [0] * NBS
        self.tma_inflight, self.mma_queue, self.mma_done, self.epilogue_started = [], [], 0, False

    def wait(self, bar, parity):
        while not bar.passed(parity):
            yield

    def tma_engine(self):
        while True:
            if self.tma_inflight:
                kind, st, c = self.tma_inflight.pop(self.rng.randrange(len(self.tma_inflight)))
                if kind == "raw":
                    assert self.busy[st] == 0 and self.lo_pending[st] == 0, f"stage {st} overwritten while in use"
                    self.raw[st], self.lo_pending[st] = c, PW
                    self.raw_full[st].arrive()
                else:
                    assert self.B_busy[st] == 0, f"B stage {st} overwritten under a running MMA"
                    self.B[st] = c
                    self.b_full[st].arrive()
            yield

    def tensor_engine(self):
        while True:
            if self.mma_queue:
                c, r, sb, commits = self.mma_queue.pop(0)
                assert self.raw[r] == c and all(x == c for x in self.lo[r]), f"MMA of chunk {c} read stage {r}: hi {self.raw[r]} lo {self.lo[r]}"
                assert self.B[sb] == c, f"MMA of chunk {c} read B stage {sb} = {self.B[sb]}"
                self.busy[r] -= 1; self.B_busy[sb] -= 1; self.mma_done += 1
                for bar in commits:
                    bar.arrive()
            yield

    def a_tma(self):
        for c in range(self.n):
            r = c % self.NST
            yield from self.wait(self.st_empty[r], ((c // self.NST) & 1) ^ 1)
            self.raw_full[r].arrive()
            self.tma_inflight.append(("raw", r, c))
            yield

    def b_tma(self):
        for c in range(self.n):
            s = c % NBS
            yield from self.wait(self.b_empty[s], ((c // NBS) & 1) ^ 1)
            self.b_full[s].arrive()
            self.tma_inflight.append(("b", s, c))
            yield

    def lo_warp(self, w):
        for c in range(self.n):
            r = c % self.NST
            yield from self.wait(self.raw_full[r], (c // self.NST) & 1)
            assert self.raw[r] == c, f"lo warp {w} read stage {r} = {self.raw[r]} for chunk {c}"
            assert self.busy[r] == 0, f"lo warp {w} rewrote stage {r} under a running MMA"
            self.lo[r][w] = c
            self.lo_pending[r] -= 1
            yield
            self.lo_full[r].arrive()
            yield
        yield from self.wait(self.acc_full, 0)
        assert self.mma_done == self.n
        self.epilogue_started = True

    def mma(self):
        for c in range(self.n):
            r, sb = c % self.NST, c % NBS
            yield from self.wait(self.b_full[sb], (c // NBS) & 1)
            yield from self.wait(self.lo_full[r], (c // self.NST) & 1)
            commits = [self.st_empty[r], self.b_empty[sb]] + ([self.acc_full] if c == self.n - 1 else [])
            self.busy[r] += 1; self.B_busy[sb] += 1
            self.mma_queue.append((c, r, sb, commits))
            yield

    def run(self):
        agents = {"a_tma": self.a_tma(), "b_tma": self.b_tma(), "mma": self.mma(), "tma_engine": self.tma_engine(), "tensor_engine": self.tensor_engine()}
        for w in range(PW):
            agents[f"t{w}"] = self.lo_warp(w)
        finite = {k for k in agents if not k.endswith("engine")}
        speed = {k: 10.0 ** self.rng.uniform(-3, 0) for k in agents}
        steps = 0
        while finite:
            names = list(agents)
            name = self.rng.choices(names, weights=[speed[k] for k in names])[0]
            try:
                next(agents[name])
            except StopIteration:
                del agents[name]; finite.discard(name)
            steps += 1
            assert steps < 400_000_000, f"no termination (deadlock?) with {sorted(finite)} still running"
        assert self.epilogue_started


def main():
    runs = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    rng = random.Random(0)
    for pipelined in (True, False):
        for n_chunks in (1, 2, 3, 4, 5, 7, 8, 9, 16, 24):
            for _ in range(runs):
                Sim(n_chunks, pipelined, rng).run()
        print(f"{'pipelined' if pipelined else 'current  '} transform loop: {runs} random schedules x 10 chunk counts: no violation")
    for n_chunks in (1, 2, 3, 4, 5, 7, 8, 9, 16, 24):
        for _ in range(runs):
            SimSS(n_chunks, rng).run()
    print(f"shared-memory variant (apply_gemm3x_ss.cu): {runs} random schedules x 10 chunk counts: no violation")


if __name__ == "__main__":
    main()
