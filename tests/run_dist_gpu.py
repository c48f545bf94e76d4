"""Multi-GPU check (launch under torchrun on N >= 2 GPUs of one node):
  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/run_dist_gpu.py
Track-sharded fingerprinting -> hash-prefix-sharded index (NCCL all-to-all) -> routed queries with the
exact vote over the exchanged keys — and the same pass with the exchange fused into the scatter kernel over NVLink peer
memory; every rank's results must equal a single-GPU index built from all rows."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import sia_oracle as O                      # noqa: E402  (checker / data generator only)
from shazam_b200.database import FingerprintIndex       # noqa: E402
from shazam_b200.distributed import CudaShard, ShardedIndex, TrackShardedIndex, shard_tracks   # noqa: E402
from shazam_b200.fingerprinter import Fingerprinter     # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    fs, ntracks = 44100, 24
    tracks = [O.synth_track(500 + i, 12 * fs) for i in range(ntracks)]
    fp = Fingerprinter(local, max_chunk_frames=8192)
    mine = shard_tracks(ntracks, rank, world)
    b = fp.fingerprint_tracks([tracks[i] for i in mine], fan_value=15)
    songs = torch.cat([torch.full((int(b.starts[k + 1] - b.starts[k]),), int(mine[k]) + 1, dtype=torch.int32) for k in range(len(mine))])
    sharded = ShardedIndex(CudaShard(local, 1 << 22))
    sharded.insert(songs.to(dev), torch.from_numpy(b.hash).to(dev), torch.from_numpy(b.t1).to(dev))
    total = sharded.finalize()
    by_track = TrackShardedIndex(CudaShard(local, 1 << 22))
    by_track.insert(songs.to(dev), torch.from_numpy(b.hash).to(dev), torch.from_numpy(b.t1).to(dev))
    assert by_track.finalize() == total
    # reference: one index with every track's rows (fingerprinted locally, deterministic)
    allb = fp.fingerprint_tracks(tracks, fan_value=15)
    single = FingerprintIndex(local, 1 << 22)
    for i in range(ntracks):
        h, t = allb.track(i)
        single.insert(i + 1, h, t)
    assert single.finalize() == total, (single.rows, total)
    # queries: 5 s clips, round-robin by rank
    rng = np.random.default_rng(1)
    clips = [tracks[i % ntracks][(s := int(rng.integers(0, 6)) * fs): s + 5 * fs] for i in range(16)]
    myq = list(range(rank, len(clips), world))
    qb = fp.fingerprint_tracks([clips[i] for i in myq], fan_value=15)
    D, Oq = torch.from_numpy(qb.hash).to(dev), torch.from_numpy(qb.t1).to(dev)
    want = single.query_batch(D, Oq, qb.starts, 3)
    peer = ShardedIndex(sharded.backend, exchange="peer")          # the same shards, vote tuples written through NVLink
    peer._max_song, peer.entry_cap = sharded._max_song, sharded.entry_cap
    for name in ("hash", "hash small passes", "track", "peer", "peer small passes", "peer, forced fallback"):
        if name == "track":
            got = by_track.query(D, Oq, qb.starts, 3)
        elif name.startswith("peer"):
            if "fallback" in name:
                os.environ["SIA_PVOTE_CAP"] = "64"                 # every region overflows: the key exchange takes the pass
            got = peer.query(D, Oq, qb.starts, 3, queries_per_pass=3 if "small" in name else 4096)
            os.environ.pop("SIA_PVOTE_CAP", None)
        else:
            got = sharded.query(D, Oq, qb.starts, 3, queries_per_pass=4096 if name == "hash" else 3)
        for a, w in zip(got, want):
            assert torch.equal(a, w), (name, rank, a, w)
    assert peer.peer_fallbacks >= 1
    peer.close_peers()
    top = got[0][:, 0].cpu().tolist()
    assert top == [(i % ntracks) + 1 for i in myq], (top, myq)
    dist.barrier()
    if rank == 0:
        print(f"dist ok: world={world} rows={total} queries={len(clips)} identical to the single-GPU index")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
