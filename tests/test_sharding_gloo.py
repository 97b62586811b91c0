"""world_size-2 gloo test (CPU) of the multi-GPU host logic: round-robin layer sharding,
packed layout, ONE all-gather, reassembly.  The per-shard edit is done by the oracle here (test
infrastructure) — on GPUs the same code path runs the CUDA solver per rank (bench.py "sharded")."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import uce_oracle as O
        from uce_b200.sharding import all_gather_layers, shard_layers
        from uce_b200.synthetic import concept_rows, weights
        K, dims = 32, [8, 24, 16, 8, 40]
        rows = concept_rows(9, K, seed=1)
        W = weights(dims, K, seed=1)
        mine = shard_layers(len(W), world, rank)
        local = {}
        for i in mine:
            e = O.erase_exact_f64([W[i]], rows[:2], rows[7:9], rows[2:7], 1.0, 1.0, 0.5)[0]
            local[i] = torch.from_numpy(e).float()
        full = all_gather_layers(local, dims, K, torch.device("cpu"))
        ref = [torch.from_numpy(e).float() for e in O.erase_exact_f64(W, rows[:2], rows[7:9], rows[2:7], 1.0, 1.0, 0.5)]
        ok = all(torch.equal(a, b) for a, b in zip(full, ref)) and len(full) == len(dims)
        # the in-place form the drivers use: results written straight into slices of the gather buffer, one in-place all-gather
        from uce_b200.sharding import GatherPlan
        plan = GatherPlan(dims, K, world, rank, torch.device("cpu"))
        for i, v in plan.views_mine().items():
            v.copy_(local[i])
        full2 = plan.gather()
        ok = ok and all(torch.equal(a, b) for a, b in zip(full2, ref)) and sorted(plan.views_mine()) == mine
        # chunked form (large edits: a chunk's all-gather overlaps the next chunk's kernels): every chunk count gives the same result,
        # also with asynchronous collectives and with chunks in which this rank owns nothing
        for chunks in (2, 3, 7):
            planc = GatherPlan(dims, K, world, rank, torch.device("cpu"), chunks=chunks)
            views = planc.views_mine()
            works = []
            seen = []
            for c in range(planc.chunks):
                for l in planc.layers_of_chunk_mine(c):
                    views[l].copy_(local[l]); seen.append(l)
                works.append(planc.gather_chunk(c, async_op=True))
            for w in works:
                w.wait()
            fullc = [planc.view(l) for l in range(len(dims))]
            ok = ok and sorted(seen) == mine and all(torch.equal(a, b) for a, b in zip(fullc, ref))
        # sharded_edit(): the driver's loop — first non-empty chunk through the one-call edit, the others through the apply alone, one
        # asynchronous all-gather per chunk — with a stand-in solver that writes the oracle's result into the buffers it is handed
        from uce_b200.sharding import sharded_edit

        class Solver:
            def __init__(self):
                self.calls = []
            def _fill(self, w_old, w_new):
                for wo, wn in zip(w_old, w_new):
                    i = next(k for k, w in enumerate(W) if w is wo)
                    wn.copy_(ref[i])
            def edit(self, C, G, scales, n_edit, lamb, w_old, w_new, check=True):
                self.calls.append(("edit", len(w_old))); self._fill(w_old, w_new)
            def apply(self, w_old, w_new):
                self.calls.append(("apply", len(w_old))); self._fill(w_old, w_new)
            def check(self):
                self.calls.append(("check", 0))

        for chunks in (1, 2, 3):
            plans = GatherPlan(dims, K, world, rank, torch.device("cpu"), chunks=chunks)
            sv = Solver()
            fulls = sharded_edit(sv, plans, None, None, [], 2, 0.5, dict(enumerate(W)))
            kinds = [c[0] for c in sv.calls]
            ok = ok and all(torch.equal(a, b) for a, b in zip(fulls, ref))
            ok = ok and kinds.count("edit") == 1 and kinds[0] == "edit" and kinds[-1] == "check" and sum(c[1] for c in sv.calls) == len(mine)
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_gather_plan_layout_is_a_partition():
    """Every projection gets its own region of the packed buffer (no overlap, all inside), for any world size and chunk count; the
    chunk regions are contiguous, in order, and hold one equal part per rank."""
    from uce_b200.sharding import GatherPlan, default_chunks
    from uce_b200.synthetic import SD14_DIMS, SDXL_DIMS
    for dims, K in (([8, 24, 16, 8, 40], 32), (SD14_DIMS, 768), (SDXL_DIMS, 2048), ([5], 4), ([], 4)):
        for world in (1, 2, 3, 8):
            for chunks in (1, 2, 5, 64):
                plan = GatherPlan(dims, K, world, 0, torch.device("meta"), chunks=chunks)
                spans = sorted((plan.where[l][1], plan.where[l][1] + d * K) for l, d in enumerate(dims))
                assert all(a1 <= b0 for (_, a1), (b0, _) in zip(spans, spans[1:])), (dims[:3], world, chunks)
                assert not spans or (spans[0][0] >= 0 and spans[-1][1] <= plan.buf.numel())
                assert plan.base[-1] == plan.buf.numel() == world * sum(plan.part)
                mine = [l for r in range(world) for l in GatherPlan(dims, K, world, r, torch.device("meta"), chunks=chunks).views_mine()]
                assert sorted(mine) == list(range(len(dims)))
    assert default_chunks(SD14_DIMS, 768, 2) == 1 and default_chunks(SDXL_DIMS, 2048, 2) == 4 and default_chunks(SDXL_DIMS, 2048, 8) == 2


def test_layer_sharding_allgather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
