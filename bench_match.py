#!/usr/bin/env python
"""bench_match.py — match queries per second on a large synthetic index (BASELINE.json metric M2,
configs[3]/[4]).  Secondary benchmark; `bench.py` carries the headline (M1).

Index: `--tracks` synthetic tracks x ~`--rows-per-track` fingerprints.  Rows are NOT fingerprinted
from audio (8e9 rows of audio would take hours to synthesise): every track gets a random
time-ordered peak list (bins skewed towards low frequencies) and the rows come out of the real
K3 kernel — sha1("f1|f2|dt") of real peak pairs — so the key distribution has the skew of the
8.4e8-value pre-image space.  Queries are 5 s windows (106 frames) of indexed tracks with 30 % of
the peaks dropped and as many random peaks added; the expected answer (song, window start) is
known, so accuracy is checked.

N GPUs: the index is hash-prefix sharded over the ranks (ShardedIndex: NCCL all-to-all of rows at
build, of query hashes and of partial vote bins at query time, exact merge); queries are submitted
round-robin by rank.  `value` = queries per second over all ranks, inputs resident in HBM.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAMES_PER_TRACK = 3874          # 3 min @ 44.1 kHz
CLIP_FRAMES = 106                # 5 s
FAN = 15


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--tracks", type=int, default=12500, help="tracks in the whole index")
    ap.add_argument("--rows-per-track", type=int, default=80000)
    ap.add_argument("--queries", type=int, default=10000, help="queries per step over all ranks")
    ap.add_argument("--topn", type=int, default=3)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--gen-batch", type=int, default=500, help="tracks generated per batch")
    ap.add_argument("--mode", default="hash", choices=["hash", "bins", "track"],
                    help="N>1: hash-prefix sharding exchanging vote keys (default) or sorted bins; or track sharding")
    ap.add_argument("--vote-sweep", default="", help="world 1: comma list of vote settings timed after the main run, "
                    "e.g. 'sort,hash:262144,hash:4194304' (SIA_VOTE / SIA_VOTE_GROUP_TUPLES)")
    ap.add_argument("--cpu-baseline-tracks", type=int, default=0, help=">0: time the oracle port on a small index")
    return ap.parse_args()


def batch_peaks(dev, batch_id: int, n_tracks: int, peaks_per_track: int):
    """Deterministic peak lists of the tracks of one generation batch: (t[B,P], f[B,P]) sorted by (t, f)."""
    import torch
    g = torch.Generator(device=dev)
    g.manual_seed(77_000 + batch_id)
    t = torch.randint(0, FRAMES_PER_TRACK, (n_tracks, peaks_per_track), device=dev, generator=g)
    u = torch.rand((n_tracks, peaks_per_track), device=dev, generator=g)
    f = torch.clamp((2049 * u * u).long(), max=2048)
    key, _ = torch.sort(t * 4096 + f, dim=1)
    return (key // 4096).to(torch.int32), (key % 4096).to(torch.int32)


def main():
    args = parse()
    import torch
    import torch.distributed as dist
    from shazam_b200.database import FingerprintIndex
    from shazam_b200.distributed import CudaShard, ShardedIndex, TrackShardedIndex
    from shazam_b200.fingerprinter import Fingerprinter

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    P = max(2, args.rows_per_track // (FAN - 1))
    n_batches = -(-args.tracks // args.gen_batch)
    fp = Fingerprinter(local, max_chunk_frames=4096)
    fp_cap_peaks = args.gen_batch * P
    # K3 workspace is sized by the context's peak capacity: use a context big enough for one batch
    fp.close()
    os.environ["SIA_PEAKS_PER_FRAME_CAP"] = str(max(32, -(-fp_cap_peaks // 4096) + 1))
    fp = Fingerprinter(local, max_chunk_frames=4096)

    rows_total_est = args.tracks * args.rows_per_track
    cap = int(rows_total_est / world * 1.15) + (1 << 20)
    shard = CudaShard(local, cap)
    index = (TrackShardedIndex(shard, rank=rank, world=world) if args.mode == "track" else
             ShardedIndex(shard, rank=rank, world=world, exchange="bins" if args.mode == "bins" else "tuples"))

    # ---- build ---------------------------------------------------------------------------------
    torch.cuda.synchronize()
    t_build = time.perf_counter()
    gen_rows = 0
    rounds = -(-n_batches // world)
    for rnd in range(rounds):
        b = rnd * world + rank
        if b < n_batches:
            nt = min(args.gen_batch, args.tracks - b * args.gen_batch)
            pt, pf = batch_peaks(dev, b, nt, P)
            tps = torch.arange(nt + 1, device=dev, dtype=torch.int64) * P
            h, t1, ths = fp.pairs_sha1(pt.reshape(-1), pf.reshape(-1), tps, FAN)
            songs = torch.repeat_interleave(torch.arange(nt, device=dev, dtype=torch.int32) + (b * args.gen_batch + 1),
                                            (ths[1:] - ths[:-1]))
            gen_rows += h.shape[0]
        else:
            h = torch.empty((0, 10), dtype=torch.uint8, device=dev)
            t1 = torch.empty(0, dtype=torch.int32, device=dev)
            songs = torch.empty(0, dtype=torch.int32, device=dev)
        index.insert(songs, h, t1)          # collective: rows travel to their owning shard
        del h, t1, songs
    rows = index.finalize()
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t_build

    # ---- queries -----------------------------------------------------------------------------------
    rng = np.random.default_rng(5)
    q_tracks = rng.integers(0, args.tracks, args.queries)
    q_start = rng.integers(0, FRAMES_PER_TRACK - CLIP_FRAMES, args.queries)
    mine = np.arange(rank, args.queries, world)
    by_batch = {}
    for qi in mine:
        by_batch.setdefault(int(q_tracks[qi]) // args.gen_batch, []).append(int(qi))
    q_pt, q_pf, q_len, order = [], [], [], []
    gq = torch.Generator(device=dev)
    gq.manual_seed(1234 + rank)
    for b, qis in sorted(by_batch.items()):
        nt = min(args.gen_batch, args.tracks - b * args.gen_batch)
        pt, pf = batch_peaks(dev, b, nt, P)
        for qi in qis:
            row = int(q_tracks[qi]) - b * args.gen_batch
            t0 = int(q_start[qi])
            m = (pt[row] >= t0) & (pt[row] < t0 + CLIP_FRAMES)
            ct, cf = pt[row][m] - t0, pf[row][m]
            keep = torch.rand(ct.shape[0], device=dev, generator=gq) < 0.7
            n_noise = int((~keep).sum().item())
            nt_ = torch.randint(0, CLIP_FRAMES, (n_noise,), device=dev, generator=gq, dtype=torch.int32)
            nf_ = torch.randint(0, 2049, (n_noise,), device=dev, generator=gq, dtype=torch.int32)
            key, _ = torch.sort(torch.cat([ct[keep], nt_]).long() * 4096 + torch.cat([cf[keep], nf_]).long())
            key = torch.unique_consecutive(key)
            q_pt.append((key // 4096).to(torch.int32)); q_pf.append((key % 4096).to(torch.int32))
            q_len.append(key.shape[0]); order.append(qi)
    tps = torch.tensor(np.cumsum([0] + q_len), dtype=torch.int64, device=dev)
    if q_len:
        qh, qt1, qths = fp.pairs_sha1(torch.cat(q_pt), torch.cat(q_pf), tps, FAN)
        q_starts = qths.cpu().numpy()
    else:
        qh = torch.empty((0, 10), dtype=torch.uint8, device=dev); qt1 = torch.empty(0, dtype=torch.int32, device=dev)
        q_starts = np.zeros(1, np.int64)
    n_q_local = len(order)

    def step():
        if world == 1:      # single GPU: the fused lookup + vote entry point
            return shard.index.query_batch(qh, qt1, q_starts, args.topn)
        return index.query(qh, qt1, q_starts, args.topn)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 1)):
        res = step()
    barrier()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    evs[0].record()
    for i in range(args.steps):
        res = step()
        evs[i + 1].record()
    barrier()
    per_step = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    ms = torch.tensor([evs[0].elapsed_time(evs[-1]) / args.steps, float(np.median(per_step))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step, ms_median = float(ms[0].item()), float(ms[1].item())

    # accuracy (the answer is known by construction)
    song = res[0][:, 0].cpu().numpy() if n_q_local else np.zeros(0, np.int32)
    diff = res[1][:, 0].cpu().numpy() if n_q_local else np.zeros(0, np.int32)
    ok = np.array([song[i] == q_tracks[qi] + 1 and diff[i] == q_start[qi] for i, qi in enumerate(order)], bool)
    stats = torch.tensor([ok.sum(), n_q_local, qt1.numel(), gen_rows], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(stats)
    stats = stats.cpu().tolist()
    # per-query workload statistics from one single-shard pass (rank 0, world 1 only: needs the whole index)
    qstats = None
    if world == 1 and n_q_local:
        out = shard.index.query_batch(qh, qt1, q_starts, args.topn, want_stats=True)
        qstats = out[5]

    # ---- end to end: query hashes start in pinned host memory, results end in host memory, every step ----------
    h_qh = qh.cpu().pin_memory(); h_qt1 = qt1.cpu().pin_memory()

    def step_host():
        d_h = h_qh.to(dev, non_blocking=True); d_t = h_qt1.to(dev, non_blocking=True)
        r = shard.index.query_batch(d_h, d_t, q_starts, args.topn) if world == 1 else index.query(d_h, d_t, q_starts, args.topn)
        return [x.cpu() for x in r[:5]]
    step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res_host = step_host()
    barrier()
    e2e_s = torch.tensor([(time.perf_counter() - t0) / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    e2e = {"value": args.queries / e2e_s, "unit": "queries/s", "ms_per_step": e2e_s * 1e3,
           "h2d_bytes_per_step_per_rank": int(h_qh.numel() + 4 * h_qt1.numel()),
           "d2h_bytes_per_step_per_rank": int(sum(x.numel() * 4 for x in res_host)),
           "api": "pinned host (digest, offset) arrays -> query_batch / ShardedIndex.query -> results in host memory"}

    sweep = {}
    if world == 1 and args.vote_sweep:
        for setting in args.vote_sweep.split(","):
            parts = setting.split(":")          # mode[:group tuples[:tuples per block[:bin slots per tuple[:filter 0/1]]]]
            os.environ["SIA_VOTE"] = parts[0]
            for name, val in zip(("SIA_VOTE_GROUP_TUPLES", "SIA_VOTE_CHUNK", "SIA_VOTE_LOAD", "SIA_VOTE_FILTER"), parts[1:] + [""] * 4):
                if val:
                    os.environ[name] = val
                else:
                    os.environ.pop(name, None)
            step(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                step()
            torch.cuda.synchronize()
            sweep[setting] = round((time.perf_counter() - t0) / args.steps * 1e3, 2)
        for name in ("SIA_VOTE", "SIA_VOTE_GROUP_TUPLES", "SIA_VOTE_CHUNK", "SIA_VOTE_LOAD", "SIA_VOTE_FILTER"):
            os.environ.pop(name, None)

    if rank == 0:
        line = {
            "metric": "match_queries_per_second", "value": args.queries / (ms_step * 1e-3), "unit": "queries/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": ms_step,
            "ms_per_step_median": ms_median, "ms_per_step_each_rank0": [round(x, 2) for x in per_step],
            "higher_is_better": True, "scaling": "strong (fixed index and query set)",
            "sharding": {"hash": "hash prefix, vote keys exchanged, owner votes with hash tables (exact)",
                         "bins": "hash prefix, sorted partial bins exchanged, owner re-sorts and sums (exact)",
                         "track": "by track, queries all-gathered, G x topn candidates merged (exact)"}[args.mode],
            "dtype": "u64 keys / 16-byte rows", "data": "synthetic",
            "config": {"workload": f"{args.queries} concurrent 5 s queries (topn {args.topn}) against a {args.tracks}-track "
                                   f"index, ~{args.rows_per_track} rows per track, {args.mode}-sharded over {world} GPU(s)",
                       "index_rows": rows, "rows_generated": stats[3], "query_hashes_per_step": stats[2],
                       "mean_hashes_per_query": stats[2] / max(1, args.queries)},
            "accuracy_top1_song_and_offset": stats[0] / max(1, stats[1]),
            "index_build_seconds": round(t_build, 2),
            "index_build_rows_per_second": rows / t_build,
            "e2e": e2e,
        }
        if sweep:
            line["vote_sweep_ms_per_step"] = sweep
        if qstats:
            line["per_step"] = {"query_pairs": qstats[0], "db_rows_matched": qstats[1], "vote_tuples": qstats[2],
                                "distinct_bins": qstats[3],
                                "mean_postings_per_query_hash": qstats[1] / max(1, qstats[0])}
        if args.cpu_baseline_tracks > 0:
            line["cpu_baseline"] = cpu_baseline(args, P)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(args, P):
    """The reference's return_matches + align_matches (oracle port, in-memory table standing in for MySQL)
    on a SMALL index: the full index does not fit a Python dict (state the size with the number)."""
    import torch
    from oracle import sia_oracle as O
    from shazam_b200.fingerprinter import Fingerprinter
    ntr = args.cpu_baseline_tracks
    dev = torch.device("cuda", 0)
    fp = Fingerprinter(0, max_chunk_frames=4096)
    pt, pf = batch_peaks(dev, 0, ntr, P)
    tps = torch.arange(ntr + 1, device=dev, dtype=torch.int64) * P
    h, t1, ths = fp.pairs_sha1(pt.reshape(-1), pf.reshape(-1), tps, FAN)
    h = h.cpu().numpy(); t1 = t1.cpu().numpy(); ths = ths.cpu().numpy()
    table = O.FingerprintTable()
    for s in range(ntr):
        sid = table.insert_song(f"t{s}", "AB" * 20, int(ths[s + 1] - ths[s]))
        hx = h[ths[s]:ths[s + 1]].tobytes().hex()
        table.insert_hashes(sid, [(hx[20 * i:20 * i + 20], int(o)) for i, o in enumerate(t1[ths[s]:ths[s + 1]])])
    rng = np.random.default_rng(9)
    queries = []
    for _ in range(40):
        s = int(rng.integers(0, ntr)); t0 = int(rng.integers(0, FRAMES_PER_TRACK - CLIP_FRAMES))
        sl = slice(ths[s], ths[s + 1])
        m = (t1[sl] >= t0) & (t1[sl] < t0 + CLIP_FRAMES - 20)
        hx = h[sl][m].tobytes().hex()
        queries.append(set((hx[20 * i:20 * i + 20], int(o) - t0) for i, o in enumerate(t1[sl][m])))
    t0 = time.perf_counter()
    for q in queries:
        mt, dd = O.return_matches(table, q)
        O.align_matches(table, mt, dd, len(q), args.topn)
    dt = time.perf_counter() - t0
    return {"value": len(queries) / dt, "unit": "queries/s", "cores": 1, "kind": "port",
            "sample": f"{len(queries)} queries against a {ntr}-track ({table.num_rows()} rows) in-memory table; "
                      "the full index does not fit a Python dict, and postings per hash grow with index size"}


if __name__ == "__main__":
    main()
