"""CPU stand-in for one index shard (numpy/dicts, reference semantics, the slot formats of csrc/index_dist.cu) so that
the multi-rank orchestration in shazam_b200.distributed can run under gloo without a GPU.  Test infrastructure only."""
import numpy as np
import torch

from shazam_b200.distributed import ShardBackend, QID_BITS, SONG_BITS, DIFF_BITS

BIAS = 1 << 24
M24 = (1 << 24) - 1
U64 = (1 << 64) - 1


def _i64(v):
    """python int (uint64 range) -> the int64 that holds the same bits"""
    v &= U64
    return v - (1 << 64) if v >= 1 << 63 else v


def pack_entry(qid, digest: bytes, qoff):
    hi = int.from_bytes(digest[:8], "big")
    lo16 = int.from_bytes(digest[8:10], "big")
    x = ((hi & M24) << 40) | (lo16 << 24) | (qoff & M24)
    y = (qid << 40) | (hi >> 24)
    return _i64(x), _i64(y)


def unpack_entry(x, y):
    x &= U64; y &= U64
    qid = y >> 40
    hi = ((y << 24) & U64) | (x >> 40)
    lo16 = (x >> 24) & 0xffff
    return qid, hi.to_bytes(8, "big") + lo16.to_bytes(2, "big"), x & M24


class _Counters:
    def __init__(self, owner):
        self.owner = owner

    def zero_(self):
        self.owner.tuples = []


class CpuPeers:
    """Stand-in for database.PeerBuffers."""

    def __init__(self, rank, world, qp, region_cap, group):
        self.rank, self.world, self.qp, self.region_cap, self.group = rank, world, qp, region_cap, group
        self.tuples = []
        self.counters = _Counters(self)

    def close(self, group=None):
        pass


class CpuShard(ShardBackend):
    def __init__(self):
        self.device = torch.device("cpu")
        self.rows = {}       # digest bytes -> set[(song, off)]
        self.n = 0
        self._max_song = 0

    def insert_rows(self, songs, digests, offsets):
        for s, d, o in zip(songs.tolist(), digests.numpy(), offsets.tolist()):
            self.rows.setdefault(bytes(d), set()).add((int(s), int(o)))
            self._max_song = max(self._max_song, int(s))

    def finalize(self):
        self.n = sum(len(v) for v in self.rows.values())
        return self.n

    def max_song(self):
        return self._max_song

    # ---- the three device steps of a hash-prefix query pass --------------------------------------------------
    def route_entries(self, digests, qoffsets, query_starts, qid_base, world, slot_cap, status):
        slots = torch.zeros((world, slot_cap, 2), dtype=torch.int64)
        qs = query_starts.tolist()
        counts = [0] * world
        dn = digests.numpy()
        for q in range(len(qs) - 1):
            for i in range(qs[q], qs[q + 1]):
                d = bytes(dn[i])
                owner = (int.from_bytes(d[:2], "big") * world) >> 16
                counts[owner] += 1
                if counts[owner] < slot_cap:
                    x, y = pack_entry(qid_base + q, d, int(qoffsets[i]))
                    slots[owner, counts[owner], 0] = x
                    slots[owner, counts[owner], 1] = y
        for d in range(world):
            slots[d, 0, 0] = counts[d]
        return slots

    def expand_slots(self, entry_slots, world, queries_per_rank, key_cap, info):
        ent = set()
        cap = entry_slots.shape[1]
        for s in range(world):
            c = int(entry_slots[s, 0, 0])
            if c > cap - 1:
                info[0] |= 1
                info[2] = max(int(info[2]), c + 1)
            for k in range(1, min(c, cap - 1) + 1):
                ent.add(unpack_entry(int(entry_slots[s, k, 0]), int(entry_slots[s, k, 1])))
        out = torch.zeros((world, key_cap), dtype=torch.int64)
        counts = [0] * world
        seen = set()
        for q, h, qo in sorted(ent):
            dest, ql = divmod(q, queries_per_rank)
            head = (q, h) not in seen
            seen.add((q, h))
            for song, off in sorted(self.rows.get(h, ())):
                key = (int(head) << 63) | (ql << (SONG_BITS + DIFF_BITS)) | (song << DIFF_BITS) | (off - qo + BIAS)
                counts[dest] += 1
                if counts[dest] < key_cap:
                    out[dest, counts[dest]] = _i64(key)
        for d in range(world):
            out[d, 0] = counts[d]
            info[1] = max(int(info[1]), counts[d] + 1)
            if counts[d] > key_cap - 1:
                info[0] |= 2
        return out

    def vote_finish(self):
        pass

    # ---- the peer-memory pass: the "regions" of a rank are a list its peers fill through the process group ----------
    def make_peers(self, rank, world, qp, region_cap, fill_cap, group=None):
        return CpuPeers(rank, world, qp, region_cap, group)

    def lookup_slots(self, entry_slots, world, queries_per_rank, info):
        ent = set()
        cap = entry_slots.shape[1]
        for s in range(world):
            c = int(entry_slots[s, 0, 0])
            if c > cap - 1:
                info[0] |= 1
                info[2] = max(int(info[2]), c + 1)
            for k in range(1, min(c, cap - 1) + 1):
                ent.add(unpack_entry(int(entry_slots[s, k, 0]), int(entry_slots[s, k, 1])))
        self._ent = sorted(ent)
        t = torch.zeros(world * queries_per_rank, dtype=torch.int64)
        for q, h, _ in self._ent:
            t[q] += len(self.rows.get(h, ()))
        return t

    def scatter_peers(self, world, queries_per_rank, tuples_total, peers, info):
        import torch.distributed as dist
        need = max(int(tuples_total[d * queries_per_rank:(d + 1) * queries_per_rank].sum()) for d in range(world))
        info[3] = max(int(info[3]), need)
        out = [[] for _ in range(world)]
        if need > peers.region_cap:
            info[0] |= 4                                        # every rank sees the same totals: nobody scatters
        else:
            seen = set()
            for q, h, qo in self._ent:
                dest, ql = divmod(q, queries_per_rank)
                head = (q, h) not in seen
                seen.add((q, h))
                for song, off in sorted(self.rows.get(h, ())):
                    out[dest].append((int(head) << 63) | (ql << (SONG_BITS + DIFF_BITS)) | (song << DIFF_BITS) | (off - qo + BIAS))
        got = [None] * world
        dist.all_gather_object(got, out, group=peers.group)      # stands in for the stores into the owners' memory
        for src in range(world):
            peers.tuples += got[src][peers.rank]

    def count_regions(self, tuples_total, n_queries, topn, peers, info):
        import os
        cap = int(os.environ.get("SIA_PVOTE_CAP", 24576))
        bins = {}
        for k in peers.tuples:
            key = k & ((1 << 63) - 1)
            bins[key] = bins.get(key, 0) + 1
        flagged = {(key >> (SONG_BITS + DIFF_BITS)) & ((1 << QID_BITS) - 1) for key, c in bins.items() if c > cap}
        info[1] += len(flagged)                                  # a bin above the region size: the key exchange takes the pass
        keys = [k for k in peers.tuples if ((k >> (SONG_BITS + DIFF_BITS)) & ((1 << QID_BITS) - 1)) not in flagged]
        return vote_keys(keys, n_queries, topn)

    def vote_key_slots(self, key_slots, n_queries, topn, max_song, defer=False):
        keys = []
        cap = key_slots.shape[1]
        for s in range(key_slots.shape[0]):
            c = int(key_slots[s, 0])
            assert c <= cap - 1
            keys += [int(k) & U64 for k in key_slots[s, 1:c + 1].tolist()]
        return vote_keys(keys, n_queries, topn)

    def query_batch(self, digests, qoffsets, query_starts, topn):
        qs = torch.as_tensor(np.asarray(query_starts), dtype=torch.int64)
        nq = len(qs) - 1
        total = int(qs[-1])
        status = torch.zeros(1, dtype=torch.int32)
        info = torch.zeros(4, dtype=torch.int64)
        slots = self.route_entries(digests, qoffsets, qs, 0, 1, total + 2, status)
        keys = self.expand_slots(slots, 1, max(nq, 1), 1 + sum(len(v) for v in self.rows.values()) * max(1, total), info)
        return self.vote_key_slots(keys, nq, topn, self._max_song)


def vote_keys(keys, n_queries, topn):
    """align_matches over vote keys (python ints): per (query, song) the bin with the largest count (smallest diff on
    ties), per query the topn songs by (count desc, song asc); rows = head keys per (query, song)."""
    bins, rows = {}, {}
    qmask = (1 << QID_BITS) - 1
    for k in keys:
        q = (k >> (SONG_BITS + DIFF_BITS)) & qmask
        song = (k >> DIFF_BITS) & M24
        diff = (k & ((1 << DIFF_BITS) - 1)) - BIAS
        bins[(q, song, diff)] = bins.get((q, song, diff), 0) + 1
        if k >> 63:
            rows[(q, song)] = rows.get((q, song), 0) + 1
    outs = [torch.zeros((n_queries, topn), dtype=torch.int32) for _ in range(4)]
    nres = torch.zeros(n_queries, dtype=torch.int32)
    per_q = {}
    for (q, song, diff), c in bins.items():
        best = per_q.setdefault(q, {}).get(song)
        if best is None or c > best[0] or (c == best[0] and diff < best[1]):
            per_q[q][song] = (c, diff)
    for q, songs in per_q.items():
        ranked = sorted(songs.items(), key=lambda kv: (-kv[1][0], kv[0]))[:topn]
        nres[q] = len(ranked)
        for r, (song, (c, diff)) in enumerate(ranked):
            outs[0][q, r], outs[1][q, r], outs[2][q, r] = song, diff, c
            outs[3][q, r] = rows.get((q, song), 0)
    return (*outs, nres)
