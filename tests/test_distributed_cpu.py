"""world_size-2 gloo tests of the multi-GPU host logic (track sharding, hash-prefix routing through fixed-size slots,
overflow retries, the exact vote over the keys of all shards, the peer-memory pass with its flag handling, region growth
and fallback to the key exchange) with a CPU stand-in shard.  The CUDA shard is covered by tests/test_index_gpu.py."""
import hashlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sia_oracle as O


def _hx(i):
    return hashlib.sha1(str(i).encode()).digest()[:10]


def _world(seed=3, nsongs=12, per_song=120, universe=300):
    rng = np.random.default_rng(seed)
    songs = []
    for s in range(nsongs):
        rows = [(_hx(int(rng.integers(0, universe))), int(rng.integers(0, 200))) for _ in range(per_song)]
        songs.append((s + 1, rows))
    queries = []
    for qi in range(9):
        sid, rows = songs[int(rng.integers(0, nsongs))]
        q = {(h, max(0, o - 13)) for h, o in rows[:50]} | {(_hx(int(rng.integers(0, 2 * universe))), 2) for _ in range(25)}
        queries.append(sorted(q))
    queries[3] = []
    return songs, queries


def _oracle_results(songs, queries, topn):
    table = O.FingerprintTable()
    for sid, rows in songs:
        assert table.insert_song(f"s{sid}", "AB" * 20, len(rows)) == sid
        table.insert_hashes(sid, [(h.hex(), o) for h, o in rows])
    out = []
    for q in queries:
        m, dd = O.return_matches(table, [(h.hex(), o) for h, o in q])
        out.append([(s, d, c, dd[s]) for s, d, c in O.best_offsets(m, topn)])
    return out, table.num_rows()


def _worker(rank, world, port, topn, mode="hash"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from shazam_b200.distributed import ShardedIndex, TrackShardedIndex, shard_tracks, hash_owner, gather_results
        from tests.dist_helpers import CpuShard
        songs, queries = _world()
        want, nrows = _oracle_results(songs, queries, topn)
        if mode == "track":
            ix = TrackShardedIndex(CpuShard())
        elif mode == "peer":      # the second exchange fused into the shards' scatter (stand-in: lists filled through the group)
            ix = ShardedIndex(CpuShard(), key_cap=8, exchange="peer", region_cap=64)
        else:
            ix = ShardedIndex(CpuShard(), key_cap=8)
        mine = shard_tracks(len(songs), rank, world)            # tracks fingerprinted by this rank
        assert sorted(np.concatenate([shard_tracks(len(songs), r, world) for r in range(world)]).tolist()) == list(range(len(songs)))
        sid = torch.tensor([songs[i][0] for i in mine for _ in songs[i][1]], dtype=torch.int32)
        dig = torch.tensor(np.array([np.frombuffer(h, np.uint8) for i in mine for h, _ in songs[i][1]]).reshape(-1, 10))
        off = torch.tensor([o for i in mine for _, o in songs[i][1]], dtype=torch.int32)
        ix.insert(sid, dig, off)
        assert ix.finalize() == nrows                           # set semantics survive the exchange
        if mode != "track":       # every row landed on the shard that owns its hash prefix
            for h in ix.backend.rows:
                assert int(hash_owner(torch.tensor(np.frombuffer(h, np.uint8)).reshape(1, 10), world)) == rank
        else:                     # every rank kept exactly the songs it fingerprinted
            assert {s for v in ix.backend.rows.values() for s, _ in v} == {songs[i][0] for i in mine}
        # queries are submitted round-robin by rank
        myq = list(range(rank, len(queries), world))
        D = np.array([np.frombuffer(h, np.uint8) for qi in myq for h, _ in queries[qi]], np.uint8).reshape(-1, 10)
        Oq = np.array([o for qi in myq for _, o in queries[qi]], np.int32)
        starts = np.cumsum([0] + [len(queries[qi]) for qi in myq])
        kw = dict(queries_per_pass=3) if mode in ("hash", "peer") else {}      # forces split passes; tiny first slots force retries
        res = ix.query(torch.from_numpy(D), torch.from_numpy(Oq), starts, topn, **kw)
        if mode == "peer":
            assert ix.retries >= 1 and ix.peer_fallbacks == 0   # the regions were grown, no pass needed the key exchange
            os.environ["SIA_PVOTE_CAP"] = "2"                   # now every true bin is "above the region size"
            res3 = ix.query(torch.from_numpy(D), torch.from_numpy(Oq), starts, topn, **kw)
            os.environ.pop("SIA_PVOTE_CAP")
            assert ix.peer_fallbacks >= 1                       # those passes were redone with the key exchange
            for a, b in zip(res, res3):
                assert torch.equal(a, b)
        if mode in ("hash", "peer"):
            assert ix.retries >= 1                              # the first pass outgrew the initial key slots
            before = ix.retries
            res2 = ix.query(torch.from_numpy(D), torch.from_numpy(Oq), starts, topn, **kw)
            assert ix.retries == before                         # steady state: capacities are kept, no retry
            for a, b in zip(res, res2):
                assert torch.equal(a, b)
        song, diff, cnt, rows, nres = [t.numpy() for t in res]
        for k, qi in enumerate(myq):
            got = [(int(song[k, r]), int(diff[k, r]), int(cnt[k, r]), int(rows[k, r])) for r in range(nres[k])]
            assert got == want[qi], (rank, qi, got, want[qi])
        allres = gather_results(res)
        if rank == 0:
            assert allres[4].shape[0] == len(queries)
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("topn,mode", [(1, "hash"), (3, "hash"), (2, "track"), (3, "peer")])
def test_sharded_index_world2_gloo(topn, mode):
    mp.spawn(_worker, args=(2, _free_port(), topn, mode), nprocs=2, join=True)


@pytest.mark.parametrize("world,topn,mode", [(4, 3, "hash"), (4, 3, "peer"), (4, 2, "track"), (3, 2, "hash"), (3, 3, "peer")])
def test_sharded_index_other_world_sizes_gloo(world, topn, mode):
    """The same checks at world 4 (the N the GPU box runs between 2 and 8) and at an odd world size: 9 queries over 4 / 3
    ranks give the ranks different numbers of queries and passes (stub passes on the ranks that have run out)."""
    mp.spawn(_worker, args=(world, _free_port(), topn, mode), nprocs=world, join=True)


def test_single_rank_path_matches_oracle():
    """world=1 takes no collective at all."""
    from shazam_b200.distributed import ShardedIndex, hash_owner
    from tests.dist_helpers import CpuShard
    songs, queries = _world(seed=8)
    want, nrows = _oracle_results(songs, queries, 2)
    ix = ShardedIndex(CpuShard(), rank=0, world=1)
    sid = torch.tensor([s for s, rows in songs for _ in rows], dtype=torch.int32)
    dig = torch.tensor(np.array([np.frombuffer(h, np.uint8) for _, rows in songs for h, _ in rows]).reshape(-1, 10))
    off = torch.tensor([o for _, rows in songs for _, o in rows], dtype=torch.int32)
    ix.insert(sid, dig, off)
    assert ix.finalize() == nrows
    D = np.array([np.frombuffer(h, np.uint8) for q in queries for h, _ in q], np.uint8).reshape(-1, 10)
    Oq = np.array([o for q in queries for _, o in q], np.int32)
    starts = np.cumsum([0] + [len(q) for q in queries])
    song, diff, cnt, rows, nres = [t.numpy() for t in ix.query(torch.from_numpy(D), torch.from_numpy(Oq), starts, 2)]
    for qi in range(len(queries)):
        got = [(int(song[qi, r]), int(diff[qi, r]), int(cnt[qi, r]), int(rows[qi, r])) for r in range(nres[qi])]
        assert got == want[qi]
    own = hash_owner(torch.from_numpy(D), 8)
    assert own.min() >= 0 and own.max() <= 7 and len(torch.unique(own)) == 8     # prefixes spread over 8 shards
