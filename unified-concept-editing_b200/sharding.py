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
    """The packed symmetric buffer of one sharded edit: ``world`` equal slices, slice r holds the projections of rank r back to back.
    A rank hands ``views_mine()`` to the solver as its W_new tensors, so the apply kernels write the edited weights where the
    collective reads them (no zero fill, no pack copy), and ``gather()`` is ONE in-place all-gather; the results are views too
    (no unpack copy)."""

    def __init__(self, dims, K: int, world: int, rank: int, device):
        self.dims, self.K, self.world, self.rank = list(dims), int(K), int(world), int(rank)
        self.per_rank, self.where = pack_layout(self.dims, self.K, self.world)
        self.buf = torch.empty(self.world * self.per_rank, dtype=torch.float32, device=device)

    def view(self, l: int) -> torch.Tensor:
        r, off = self.where[l]
        d = self.dims[l]
        return self.buf[r * self.per_rank + off: r * self.per_rank + off + d * self.K].view(d, self.K)

    def views_mine(self):
        return {l: self.view(l) for l in shard_layers(len(self.dims), self.world, self.rank)}

    def gather(self, group=None):
        import torch.distributed as dist
        mine = self.buf[self.rank * self.per_rank: (self.rank + 1) * self.per_rank]
        dist.all_gather_into_tensor(self.buf, mine, group=group)       # in place: the input is this rank's slice of the output
        return [self.view(l) for l in range(len(self.dims))]


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
