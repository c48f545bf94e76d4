"""Multi-GPU sharding of the path: one process per GPU, ``torch.distributed`` for the plumbing.

* Fingerprinting shards BY TRACK with no collective — the reference's own data-parallel
  unit (``imap_unordered`` over files, ``__init__.py:357``): ``shard_tracks``.
* The index shards BY HASH PREFIX: ``owner = floor(prefix16(hash) * world / 65536)``.
  - build: fingerprints are produced track-sharded, so rows are exchanged once
    (``all_to_all``) to their owning shard;
  - query: exchange #1 routes every query hash (+ its query offset and query id) to the
    shard that owns it; each shard computes PARTIAL vote histograms
    ``(query, song, diff) -> count`` and ``(query, song) -> rows``.  A bin's true count is
    the SUM over shards, so exchange #2 sends the partial bins to the rank that owns the
    query, which sums equal keys and votes (``sia_vote_bins``); results can then be
    gathered.  This keeps results identical to the single-GPU index, tie-breaks included.

The exchanges are NCCL all-to-alls over NVLink/NVSwitch; the path has no other collective.
Everything here is host-side orchestration over a ``ShardBackend`` — the CUDA one wraps
``FingerprintIndex``; tests inject a CPU stand-in to exercise the routing under gloo.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

QID_BITS, SONG_BITS, DIFF_BITS = 15, 24, 25
MAX_QUERIES_PER_PASS = 1 << QID_BITS


def shard_tracks(n_tracks: int, rank: int, world: int) -> np.ndarray:
    """Round-robin track ownership (no collective on the fingerprint path)."""
    return np.arange(rank, n_tracks, world, dtype=np.int64)


def hash_owner(digests: torch.Tensor, world: int) -> torch.Tensor:
    """Owning shard of each digest (uint8[n,10]) by its 16-bit prefix; int64[n]."""
    if digests.numel() == 0:
        return torch.empty(0, dtype=torch.int64, device=digests.device)
    p = (digests[:, 0].to(torch.int64) << 8) | digests[:, 1].to(torch.int64)
    return (p * world) >> 16


def exchange(buffers: Sequence[torch.Tensor], dest: torch.Tensor, world: int, group=None) -> List[torch.Tensor]:
    """Send row i of every buffer to rank ``dest[i]``; returns what this rank receives,
    ordered by source rank (stable within a source)."""
    if world == 1:
        return [b for b in buffers]
    order = torch.argsort(dest, stable=True)
    counts = torch.bincount(dest, minlength=world).to(torch.int64)
    recv_counts = torch.empty_like(counts)
    dist.all_to_all_single(recv_counts, counts, group=group)
    in_split = counts.tolist()
    out_split = recv_counts.tolist()
    total = int(sum(out_split))
    out = []
    for b in buffers:
        send = b[order].contiguous()
        recv = b.new_empty((total,) + tuple(b.shape[1:]))
        dist.all_to_all_single(recv, send, output_split_sizes=out_split, input_split_sizes=in_split, group=group)
        out.append(recv)
    return out


class ShardBackend:
    """What the orchestration needs from one shard."""

    device: torch.device

    def insert_rows(self, songs: torch.Tensor, digests: torch.Tensor, offsets: torch.Tensor) -> None:
        raise NotImplementedError

    def finalize(self) -> int:
        raise NotImplementedError

    def query_partial(self, digests, qoffsets, qids) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        raise NotImplementedError

    def vote(self, bin_key, bin_count, row_key, row_count, n_queries: int, topn: int):
        raise NotImplementedError


class CudaShard(ShardBackend):
    """One ``sia_index`` on this rank's GPU."""

    def __init__(self, device: int, capacity_rows: int):
        from .database import FingerprintIndex
        self.index = FingerprintIndex(device, capacity_rows)
        self.device = self.index.tdev
        self._dev_index = device

    def insert_rows(self, songs, digests, offsets):
        self.index.insert_rows(songs, digests, offsets)

    def finalize(self) -> int:
        return self.index.finalize()

    def query_partial(self, digests, qoffsets, qids):
        return self.index.query_partial(digests, qoffsets, qids)

    def vote(self, bin_key, bin_count, row_key, row_count, n_queries, topn):
        from .database import vote_bins
        return vote_bins(self._dev_index, bin_key, bin_count, row_key, row_count, n_queries, topn)

    def close(self):
        self.index.close()


class ShardedIndex:
    """The hash-prefix-sharded fingerprints table over ``world`` ranks."""

    def __init__(self, backend: ShardBackend, rank: Optional[int] = None, world: Optional[int] = None, group=None):
        self.backend = backend
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world

    # ---- build ---------------------------------------------------------------------------
    def insert(self, songs: torch.Tensor, digests: torch.Tensor, offsets: torch.Tensor) -> None:
        """Collective.  Each rank passes the rows IT produced (its tracks); rows travel to the
        shard owning their hash.  ``songs`` are global song ids."""
        dest = hash_owner(digests, self.world)
        s, d, o = exchange([songs.to(torch.int32), digests, offsets.to(torch.int32)], dest, self.world, self.group)
        self.backend.insert_rows(s, d, o)

    def finalize(self) -> int:
        """Collective.  Returns the total number of stored rows over all shards."""
        n = torch.tensor([self.backend.finalize()], dtype=torch.int64, device=self.backend.device)
        if self.world > 1:
            dist.all_reduce(n, group=self.group)
        return int(n.item())

    # ---- query ---------------------------------------------------------------------------
    def query(self, digests: torch.Tensor, qoffsets: torch.Tensor, query_starts: np.ndarray, topn: int):
        """Collective.  Each rank submits ITS queries (``query_starts`` local, int64[Q_r+1]) and gets
        their results back: int32 tensors (song[Q_r,topn], diff, count, rows, nres[Q_r])."""
        dev = self.backend.device
        qs = np.asarray(query_starts, np.int64)
        q_local = len(qs) - 1
        counts = torch.tensor([q_local], dtype=torch.int64, device=dev)
        if self.world > 1:
            allc = [torch.zeros_like(counts) for _ in range(self.world)]
            dist.all_gather(allc, counts, group=self.group)
            per_rank = [int(c.item()) for c in allc]
        else:
            per_rank = [q_local]
        max_q = max(per_rank) if per_rank else 0
        outs = [torch.zeros((q_local, topn), dtype=torch.int32, device=dev) for _ in range(4)]
        nres = torch.zeros(q_local, dtype=torch.int32, device=dev)
        # passes of at most 32768 queries in total: every rank contributes the same local range per pass
        step = max(1, MAX_QUERIES_PER_PASS // self.world)
        for lo in range(0, max_q, step):
            a, b = min(lo, q_local), min(lo + step, q_local)
            sizes = [max(0, min(lo + step, c) - min(lo, c)) for c in per_rank]
            base = int(sum(sizes[: self.rank]))
            res = self._query_pass(digests[qs[a]:qs[b]], qoffsets[qs[a]:qs[b]], qs[a:b + 1] - qs[a], sizes, base, topn)
            for o, r in zip(outs, res[:4]):
                o[a:b] = r
            nres[a:b] = res[4]
        return (*outs, nres)

    def _query_pass(self, digests, qoffsets, qs, sizes, base, topn):
        dev = self.backend.device
        nq = len(qs) - 1
        lens = torch.as_tensor(np.diff(qs), dtype=torch.int64, device=dev)
        qid = torch.repeat_interleave(torch.arange(nq, dtype=torch.int64, device=dev), lens) + base   # pass-global ids
        # exchange #1: query hashes to their owning shard
        dest = hash_owner(digests, self.world)
        d, o, q = exchange([digests, qoffsets.to(torch.int32), qid.to(torch.int32)], dest, self.world, self.group)
        bk, bc, rk, rc = self.backend.query_partial(d, o, q)
        # exchange #2: partial bins to the rank that owns the query (sum-by-key happens there)
        bounds = torch.as_tensor(np.cumsum(sizes), dtype=torch.int64, device=dev)
        shift = SONG_BITS + DIFF_BITS
        qmask = (1 << QID_BITS) - 1          # keys are uint64 carried in int64 tensors
        bk2, bc2 = exchange([bk, bc], torch.bucketize((bk >> shift) & qmask, bounds, right=True), self.world, self.group)
        rk2, rc2 = exchange([rk, rc], torch.bucketize((rk >> shift) & qmask, bounds, right=True), self.world, self.group)
        # local query ids for the vote
        bk2 = bk2 - (base << shift)
        rk2 = rk2 - (base << shift)
        return self.backend.vote(bk2, bc2, rk2, rc2, nq, topn)


def gather_results(results, group=None, dst: int = 0):
    """Collective.  Concatenate every rank's (song, diff, count, rows, nres) on rank ``dst`` in rank
    order — the top-k merge step: each query's top-n is already final on its owner."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return results
    rank = dist.get_rank(group)
    out = []
    for t in results:
        n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
        ns = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(ns, n, group=group)
        bufs = [t.new_empty((int(k.item()),) + tuple(t.shape[1:])) for k in ns]
        dist.all_gather(bufs, t.contiguous(), group=group) if len({int(k.item()) for k in ns}) == 1 else \
            _all_gather_ragged(bufs, t.contiguous(), group)
        out.append(torch.cat(bufs) if rank == dst else None)
    return out if rank == dst else None


def _all_gather_ragged(bufs, t, group):
    for src, b in enumerate(bufs):
        if src == dist.get_rank(group):
            b.copy_(t)
        dist.broadcast(b, src=dist.get_global_rank(group, src) if group is not None else src, group=group)
