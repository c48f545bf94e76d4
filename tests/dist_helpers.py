"""CPU stand-in for one index shard (numpy/dicts, reference semantics) so that the multi-rank
routing in shazam_b200.distributed can run under gloo without a GPU.  Test infrastructure only."""
import numpy as np
import torch

from shazam_b200.distributed import ShardBackend, QID_BITS, SONG_BITS, DIFF_BITS

BIAS = 1 << 24


class CpuShard(ShardBackend):
    def __init__(self):
        self.device = torch.device("cpu")
        self.rows = {}       # digest bytes -> set[(song, off)]
        self.n = 0

    def insert_rows(self, songs, digests, offsets):
        for s, d, o in zip(songs.tolist(), digests.numpy(), offsets.tolist()):
            self.rows.setdefault(bytes(d), set()).add((int(s), int(o)))

    def finalize(self):
        self.n = sum(len(v) for v in self.rows.values())
        return self.n

    def query_partial(self, digests, qoffsets, qids):
        bins, rowbins = {}, {}
        entries = set(zip(qids.tolist(), [bytes(d) for d in digests.numpy()], qoffsets.tolist()))
        heads = {(q, h) for q, h, _ in entries}
        for q, h, qo in entries:
            for song, off in self.rows.get(h, ()):
                k = (q << (SONG_BITS + DIFF_BITS)) | (song << DIFF_BITS) | (off - qo + BIAS)
                bins[k] = bins.get(k, 0) + 1
        for q, h in heads:
            for song, off in self.rows.get(h, ()):
                k = (q << (SONG_BITS + DIFF_BITS)) | (song << DIFF_BITS)
                rowbins[k] = rowbins.get(k, 0) + 1

        def pack(d):
            ks = sorted(d)
            return (torch.tensor(ks, dtype=torch.int64).reshape(-1), torch.tensor([d[k] for k in ks], dtype=torch.int32).reshape(-1))
        bk, bc = pack(bins)
        rk, rc = pack(rowbins)
        return bk, bc, rk, rc

    def expand(self, digests, qoffsets, qids, n_queries):
        entries = sorted(set(zip(qids.tolist(), [bytes(d) for d in digests.numpy()], qoffsets.tolist())))
        tk, rk = [], []
        ts, rs = [0] * (n_queries + 1), [0] * (n_queries + 1)
        seen = set()
        for q, h, qo in entries:
            for song, off in sorted(self.rows.get(h, ())):
                tk.append((q << (SONG_BITS + DIFF_BITS)) | (song << DIFF_BITS) | (off - qo + BIAS))
                ts[q + 1] += 1
                if (q, h) not in seen:
                    rk.append((q << (SONG_BITS + DIFF_BITS)) | (song << DIFF_BITS))
                    rs[q + 1] += 1
            seen.add((q, h))
        ts = np.cumsum(ts); rs = np.cumsum(rs)
        return (torch.tensor(tk, dtype=torch.int64).reshape(-1), torch.tensor(rk, dtype=torch.int64).reshape(-1),
                torch.tensor(ts, dtype=torch.int64), torch.tensor(rs, dtype=torch.int64))

    def expand_size(self, digests, qoffsets, qids, n_queries):
        return int(self.expand(digests, qoffsets, qids, n_queries)[0].numel())

    def vote_tuples(self, tuple_key, row_key, n_queries, topn):
        one = lambda k: torch.ones(k.numel(), dtype=torch.int32)
        return self.vote(tuple_key, one(tuple_key), row_key, one(row_key), n_queries, topn)

    def query_batch(self, digests, qoffsets, query_starts, topn):
        qs = np.asarray(query_starts)
        qid = torch.repeat_interleave(torch.arange(len(qs) - 1), torch.as_tensor(np.diff(qs)))
        tk, rk, _, _ = self.expand(digests, qoffsets, qid, len(qs) - 1)
        return self.vote_tuples(tk, rk, len(qs) - 1, topn)

    def vote(self, bin_key, bin_count, row_key, row_count, n_queries, topn):
        bins, rows = {}, {}
        for k, c in zip(bin_key.tolist(), bin_count.tolist()):
            bins[k] = bins.get(k, 0) + c
        for k, c in zip(row_key.tolist(), row_count.tolist()):
            rows[k >> DIFF_BITS] = rows.get(k >> DIFF_BITS, 0) + c
        outs = [torch.zeros((n_queries, topn), dtype=torch.int32) for _ in range(4)]
        nres = torch.zeros(n_queries, dtype=torch.int32)
        per_q = {}
        for k, c in bins.items():
            q = k >> (SONG_BITS + DIFF_BITS)
            song = (k >> DIFF_BITS) & ((1 << SONG_BITS) - 1)
            diff = (k & ((1 << DIFF_BITS) - 1)) - BIAS
            best = per_q.setdefault(q, {}).get(song)
            if best is None or c > best[0] or (c == best[0] and diff < best[1]):
                per_q[q][song] = (c, diff)
        for q, songs in per_q.items():
            ranked = sorted(songs.items(), key=lambda kv: (-kv[1][0], kv[0]))[:topn]
            nres[q] = len(ranked)
            for r, (song, (c, diff)) in enumerate(ranked):
                outs[0][q, r], outs[1][q, r], outs[2][q, r] = song, diff, c
                outs[3][q, r] = rows.get((q << SONG_BITS) | song, 0)
        return (*outs, nres)
