"""Multi-GPU layout of the edit solve: projections are independent given the shared factor
(the loop at trainscripts/uce_sd_erase.py:56 carries no state), so they are dealt round-robin
to ranks; every rank recomputes the tiny factor; ONE all-gather of the packed edited weights
ends the solve."""
from __future__ import annotations

import torch


def shard_layers(n_layers: int, world: int, rank: int):
    """Indices of the projections owned by ``rank`` (round-robin in named_modules order)."""
    return list(range(rank, n_layers, world))


def pack_layout(dims, K: int, world: int):
    """Per-rank element offsets of the packed buffer.  Returns (per_rank_elems, {layer: (rank, offset)})."""
    fill = [0] * world
    where = {}
    for l, d in enumerate(dims):
        r = l % world
        where[l] = (r, fill[r])
        fill[r] += d * K
    per_rank = max(fill) if fill else 0
    return per_rank, where


class GatherPlan:
    """The packed symmetric buffer of one sharded edit.  With ``chunks == 1``: ``world`` equal slices, slice r holds the projections of
    rank r back to back, and ``gather()`` is ONE in-place all-gather.  With ``chunks > 1`` the buffer is a sequence of such layouts, one
    per chunk of every rank's projection list (chunk c holds the projections with local index in [jb[c], jb[c + 1]) of EVERY rank):
    ``gather_chunk(c, async_op=True)`` ships a chunk while the apply kernels work on the next one.  A rank hands ``views_mine()`` to the
    solver as its W_new tensors, so the kernels write the edited weights where the collective reads them (no zero fill, no pack copy);
    the results are views too (no unpack copy)."""

    def __init__(self, dims, K: int, world: int, rank: int, device, chunks: int = 1):
        self.dims, self.K, self.world, self.rank = list(dims), int(K), int(world), int(rank)
        L = len(self.dims)
        per_rank_count = -(-L // self.world) if L else 0
        self.chunks = max(1, min(int(chunks), max(1, per_rank_count)))
        self.jb = [round(c * per_rank_count / self.chunks) for c in range(self.chunks + 1)]      # local-index boundaries of the chunks
        chunk_of_j = [0] * per_rank_count
        for c in range(self.chunks):
            for j in range(self.jb[c], self.jb[c + 1]):
                chunk_of_j[j] = c
        fill = [[0] * self.world for _ in range(self.chunks)]
        local = {}                                    # layer -> (chunk, rank, offset inside the rank's part of the chunk)
        for l, d in enumerate(self.dims):
            r, j = l % self.world, l // self.world
            c = chunk_of_j[j]
            local[l] = (c, r, fill[c][r])
            fill[c][r] += d * self.K
        self.part = [max(f) if f else 0 for f in fill]           # elements per rank in chunk c (padded to the largest rank)
        self.base = [0]
        for c in range(self.chunks):
            self.base.append(self.base[-1] + self.world * self.part[c])
        self.per_rank = sum(self.part)
        self.where = {l: (r, self.base[c] + r * self.part[c] + off) for l, (c, r, off) in local.items()}
        self._chunk_of = {l: c for l, (c, _, _) in local.items()}
        self.buf = torch.empty(self.base[-1], dtype=torch.float32, device=device)

    def view(self, l: int) -> torch.Tensor:
        _, start = self.where[l]
        d = self.dims[l]
        return self.buf[start: start + d * self.K].view(d, self.K)

    def views_mine(self):
        return {l: self.view(l) for l in shard_layers(len(self.dims), self.world, self.rank)}

    def layers_of_chunk_mine(self, c: int):
        return [l for l in shard_layers(len(self.dims), self.world, self.rank) if self._chunk_of[l] == c]

    def gather_chunk(self, c: int, group=None, async_op: bool = False):
        """In-place all-gather of chunk c (every rank calls it, also a rank with nothing in the chunk)."""
        import torch.distributed as dist
        out = self.buf[self.base[c]: self.base[c + 1]]
        mine = out[self.rank * self.part[c]: (self.rank + 1) * self.part[c]]
        return dist.all_gather_into_tensor(out, mine, group=group, async_op=async_op)

    def gather(self, group=None):
        for c in range(self.chunks):
            self.gather_chunk(c, group=group)
        return [self.view(l) for l in range(len(self.dims))]


def sharded_edit(solver, plan: GatherPlan, C, G, scales, n_edit, lamb, w_old, group=None, check=True):
    """This rank's share of one edit, written into ``plan``'s buffer, and the gather.  ``w_old``: {layer: weight} (at least this rank's
    layers).  One chunk: factor + apply in one call, then one all-gather.  Several chunks: the first non-empty chunk goes through the
    one-call edit (the apply's first kernel runs beside the factor), the others through the apply alone, and each chunk's all-gather
    is issued asynchronously right behind its apply — NVLink carries chunk c while the tensor cores work on chunk c + 1."""
    views = plan.views_mine()
    works, factored = [], False
    for c in range(plan.chunks):
        ls = plan.layers_of_chunk_mine(c)
        if ls:
            if not factored:
                solver.edit(C, G, scales, n_edit, lamb, [w_old[l] for l in ls], [views[l] for l in ls], check=False)
                factored = True
            else:
                solver.apply([w_old[l] for l in ls], [views[l] for l in ls])
        if plan.chunks == 1:
            plan.gather_chunk(c, group=group)
        else:
            works.append(plan.gather_chunk(c, group=group, async_op=True))
    for w in works:
        w.wait()
    if check and factored:
        solver.check()
    return [plan.view(l) for l in range(len(plan.dims))]


def default_chunks(dims, K: int, world: int) -> int:
    """Chunks worth their launches: one per ~64 MB of a rank's edited weights, at most 4 (SD-1.4: 1; SDXL on 2 GPUs: 4 — measured 4.92 / 4.81 / 4.74 / 4.85 / 5.19 ms per edit with 8 / 4 / 3 / 2 / 1 chunks);
    UCE_SHARD_CHUNKS overrides."""
    import os
    if os.environ.get("UCE_SHARD_CHUNKS"):
        return max(1, int(os.environ["UCE_SHARD_CHUNKS"]))
    mine = 4.0 * K * sum(dims) / max(1, world)
    return int(max(1, min(4, mine // (64 << 20))))


def all_gather_layers(local: dict, dims, K: int, device, group=None):
    """local: {layer index: edited [d,K] tensor} for this rank's shard.  Returns the full list of
    edited weights on every rank via a single all_gather_into_tensor of padded per-rank shards."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per_rank, where = pack_layout(dims, K, world)
    send = torch.zeros(per_rank, dtype=torch.float32, device=device)
    for l, w in local.items():
        r, off = where[l]
        assert r == rank
        send[off:off + w.numel()].copy_(w.reshape(-1))
    recv = torch.empty(world * per_rank, dtype=torch.float32, device=device)
    dist.all_gather_into_tensor(recv, send, group=group)
    out = []
    for l, d in enumerate(dims):
        r, off = where[l]
        out.append(recv[r * per_rank + off: r * per_rank + off + d * K].view(d, K))
    return out
