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

import os
import sys
import time
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


def exchange_grouped(buffers: Sequence[torch.Tensor], send_counts: Sequence[int], world: int, group=None):
    """Like ``exchange`` for rows that are ALREADY grouped by destination rank (``send_counts[r]`` rows for
    rank r, in order): no reordering, just the all-to-all."""
    if world == 1:
        return [b for b in buffers]
    dev = buffers[0].device
    counts = torch.tensor(list(send_counts), dtype=torch.int64, device=dev)
    recv_counts = torch.empty_like(counts)
    dist.all_to_all_single(recv_counts, counts, group=group)
    out_split = recv_counts.tolist()
    total = int(sum(out_split))
    out = []
    for b in buffers:
        recv = b.new_empty((total,) + tuple(b.shape[1:]))
        dist.all_to_all_single(recv, b.contiguous(), output_split_sizes=out_split, input_split_sizes=list(send_counts),
                               group=group)
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

    def expand(self, digests, qoffsets, qids, n_queries: int):
        raise NotImplementedError

    def expand_size(self, digests, qoffsets, qids, n_queries: int) -> int:
        """Number of vote keys ``expand`` would produce (sizing only)."""
        raise NotImplementedError

    def vote_tuples(self, tuple_key, row_key, n_queries: int, topn: int):
        raise NotImplementedError

    def query_batch(self, digests, qoffsets, query_starts, topn: int):
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

    def expand(self, digests, qoffsets, qids, n_queries):
        return self.index.expand(digests, qoffsets, qids, n_queries)

    def expand_size(self, digests, qoffsets, qids, n_queries):
        return self.index.expand_size(digests, qoffsets, qids, n_queries)

    def vote_tuples(self, tuple_key, row_key, n_queries, topn):
        from .database import vote_tuples
        return vote_tuples(self._dev_index, tuple_key, row_key, n_queries, topn)

    def query_batch(self, digests, qoffsets, query_starts, topn):
        return self.index.query_batch(digests, qoffsets, query_starts, topn)

    def close(self):
        self.index.close()


class ShardedIndex:
    """The hash-prefix-sharded fingerprints table over ``world`` ranks."""

    def __init__(self, backend: ShardBackend, rank: Optional[int] = None, world: Optional[int] = None, group=None,
                 exchange: str = "tuples"):
        """``exchange``: what travels to the query's owner in the second all-to-all — ``"tuples"`` (unsorted
        vote keys; the owner sorts once: the default) or ``"bins"`` (each shard sorts and run-length counts
        first; the owner re-sorts and sums).  Both are exact."""
        assert exchange in ("tuples", "bins")
        self.exchange_mode = exchange
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
    def query(self, digests: torch.Tensor, qoffsets: torch.Tensor, query_starts: np.ndarray, topn: int,
              queries_per_pass: int = 4096, tuple_budget: int = 1_000_000_000):
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
        # (and at most queries_per_pass per rank).  A pass whose vote keys would exceed `tuple_budget` on some
        # shard is retried with half the queries — decided collectively, so all ranks stay in step.
        step = max(1, min(MAX_QUERIES_PER_PASS // self.world, int(queries_per_pass)))
        lo = 0
        while lo < max_q:
            a, b = min(lo, q_local), min(lo + step, q_local)
            sizes = [max(0, min(lo + step, c) - min(lo, c)) for c in per_rank]
            base = int(sum(sizes[: self.rank]))
            res = self._query_pass(digests[qs[a]:qs[b]], qoffsets[qs[a]:qs[b]], qs[a:b + 1] - qs[a], sizes, base, topn,
                                   tuple_budget if step > 1 else None)
            if res is None:
                step = max(1, step // 2)
                continue
            for o, r in zip(outs, res[:4]):
                o[a:b] = r
            nres[a:b] = res[4]
            lo += step
        return (*outs, nres)

    def _query_pass(self, digests, qoffsets, qs, sizes, base, topn, tuple_budget=None):
        dev = self.backend.device
        nq = len(qs) - 1
        timing = os.environ.get("SIA_DIST_TIMING") and self.rank == 0
        marks = []

        def mark(name):
            if timing:
                torch.cuda.synchronize(dev)
                marks.append((name, time.perf_counter()))
        mark("start")
        lens = torch.as_tensor(np.diff(qs), dtype=torch.int64, device=dev)
        qid = torch.repeat_interleave(torch.arange(nq, dtype=torch.int64, device=dev), lens) + base   # pass-global ids
        # exchange #1: query hashes to their owning shard
        dest = hash_owner(digests, self.world)
        d, o, q = exchange([digests, qoffsets.to(torch.int32), qid.to(torch.int32)], dest, self.world, self.group)
        mark("route hashes")
        shift = SONG_BITS + DIFF_BITS
        if self.exchange_mode == "tuples":
            total_q = int(sum(sizes))
            if tuple_budget is not None and self.world > 1:
                need = torch.tensor([self.backend.expand_size(d, o, q, total_q)], dtype=torch.int64, device=dev)
                dist.all_reduce(need, op=dist.ReduceOp.MAX, group=self.group)
                if int(need.item()) > tuple_budget:
                    return None
            mark("lookup (sizing)")
            tk, rk, ts, rs = self.backend.expand(d, o, q, total_q)
            mark("expand")
            # keys are grouped by ascending query id = by ascending owner rank: split points from the offsets
            cuts = torch.as_tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int64, device=ts.device)
            tcut = ts[cuts].cpu().numpy()
            rcut = rs[cuts].cpu().numpy()
            (tk2,) = exchange_grouped([tk], np.diff(tcut).tolist(), self.world, self.group)
            (rk2,) = exchange_grouped([rk], np.diff(rcut).tolist(), self.world, self.group)
            mark("exchange vote keys")
            tk2 = tk2 - (base << shift)
            rk2 = rk2 - (base << shift)
            mark("rebase")
            res = self.backend.vote_tuples(tk2, rk2, nq, topn)
            mark("vote")
            if timing:
                print("[sia dist] pass: %d local queries, %d vote keys out, %d in: " % (nq, tk.numel(), tk2.numel()) +
                      ", ".join("%s %.1f ms" % (n, (t - marks[i][1]) * 1e3) for i, (n, t) in enumerate(marks[1:])),
                      file=sys.stderr)
            return res
        bk, bc, rk, rc = self.backend.query_partial(d, o, q)
        # exchange #2: partial bins to the rank that owns the query (sum-by-key happens there)
        bounds = torch.as_tensor(np.cumsum(sizes), dtype=torch.int64, device=dev)
        qmask = (1 << QID_BITS) - 1          # keys are uint64 carried in int64 tensors
        bk2, bc2 = exchange([bk, bc], torch.bucketize((bk >> shift) & qmask, bounds, right=True), self.world, self.group)
        rk2, rc2 = exchange([rk, rc], torch.bucketize((rk >> shift) & qmask, bounds, right=True), self.world, self.group)
        # local query ids for the vote
        bk2 = bk2 - (base << shift)
        rk2 = rk2 - (base << shift)
        return self.backend.vote(bk2, bc2, rk2, rc2, nq, topn)


class TrackShardedIndex:
    """The alternative SURVEY.md §8e documents: every rank keeps ALL rows of ITS songs (the tracks it
    fingerprinted), so building needs no exchange and a song's vote bins are complete on one rank.  A query
    is broadcast (all-gather of its hashes), voted exactly and locally on every rank, and the G x topn
    candidates are merged by (count desc, song asc).  Exact, and the only traffic is the query hashes and
    G x topn results per query; the price is that every rank probes its directory for every query hash."""

    def __init__(self, backend: ShardBackend, rank: Optional[int] = None, world: Optional[int] = None, group=None):
        self.backend = backend
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world

    def insert(self, songs, digests, offsets) -> None:
        self.backend.insert_rows(songs.to(torch.int32), digests, offsets.to(torch.int32))

    def finalize(self) -> int:
        n = torch.tensor([self.backend.finalize()], dtype=torch.int64, device=self.backend.device)
        if self.world > 1:
            dist.all_reduce(n, group=self.group)
        return int(n.item())

    def _all_gather_ragged(self, t: torch.Tensor, sizes):
        mx = max(sizes)
        pad = t.new_zeros((mx,) + tuple(t.shape[1:]))
        pad[: t.shape[0]] = t
        bufs = [torch.empty_like(pad) for _ in range(self.world)]
        dist.all_gather(bufs, pad, group=self.group)
        return [b[:n] for b, n in zip(bufs, sizes)]

    def query(self, digests, qoffsets, query_starts, topn: int):
        dev = self.backend.device
        qs = np.asarray(query_starts, np.int64)
        if self.world == 1:
            return self.backend.query_batch(digests, qoffsets, qs, topn)
        meta = torch.tensor([len(qs) - 1, int(qs[-1])], dtype=torch.int64, device=dev)
        metas = [torch.zeros_like(meta) for _ in range(self.world)]
        dist.all_gather(metas, meta, group=self.group)
        nq = [int(m[0]) for m in metas]
        nh = [int(m[1]) for m in metas]
        all_d = self._all_gather_ragged(digests, nh)
        all_o = self._all_gather_ragged(qoffsets.to(torch.int32), nh)
        lens = self._all_gather_ragged(torch.as_tensor(np.diff(qs), dtype=torch.int64, device=dev), nq)
        starts = np.concatenate([[0], np.cumsum(torch.cat(lens).cpu().numpy())])
        song, diff, cnt, rows, nres = self.backend.query_batch(torch.cat(all_d), torch.cat(all_o), starts, topn)
        # candidates of MY queries from every rank: [G, Q_me, topn]
        qbase = int(sum(nq[: self.rank]))
        mine = slice(qbase, qbase + nq[self.rank])
        valid = torch.arange(topn, device=dev)[None, :] < nres[:, None]
        cnt = torch.where(valid, cnt, torch.full_like(cnt, -1))
        packed = torch.stack([song, diff, cnt, rows], 0).contiguous()            # [4, Q_all, topn]
        gathered = [torch.empty_like(packed) for _ in range(self.world)]
        dist.all_gather(gathered, packed, group=self.group)
        cand = torch.stack([g[:, mine, :] for g in gathered], 0)                 # [G, 4, Q_me, topn]
        cand = cand.permute(1, 2, 0, 3).reshape(4, nq[self.rank], self.world * topn)
        c_song, c_diff, c_cnt, c_rows = cand[0].long(), cand[1], cand[2].long(), cand[3]
        # rank by count descending, then ascending song id (stable sort of recognizer.py:307-310)
        key = torch.where(c_cnt >= 0, (c_cnt << 24) | ((1 << 24) - 1 - c_song), torch.full_like(c_cnt, -1))
        top = torch.topk(key, min(topn, key.shape[1]), dim=1).indices
        pick = lambda x: torch.gather(x, 1, top)
        o_cnt = pick(c_cnt)
        ok = o_cnt >= 0
        z = lambda x: torch.where(ok, pick(x).to(torch.int32), torch.zeros_like(o_cnt, dtype=torch.int32))
        return z(c_song), z(c_diff), z(c_cnt), z(c_rows), ok.sum(1).to(torch.int32)


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
