#!/usr/bin/env python
"""BASELINE.json configs[2]: recognise 5 s and 15 s clips with ADD_NOISE at SNR 0 / 10 dB against a
2,714-track GPU index (mirrors the reference's tests_csv runs, minus the loudspeaker/microphone loop).

* index: `--tracks` synthetic 3-min tracks (bench.py's generator), fingerprinted on the GPU, inserted through
  GPUDatabase (insert_song / insert_hashes_array / set_song_fingerprinted);
* queries: one clip per track at a seeded integer-second start (recognizer_test.py:534-541), mixed with
  band-limited noise scaled by get_noise_from_sound's rule RMS_n = RMS_s / 10^(SNR/20)
  (recognizer_test.py:426-435), quantised to int16;
* GPU path: Fingerprinter -> index.query_batch (return_matches + align_matches vote), top-1 checked against the
  known (song, offset);
* oracle agreement on a sample of the queries: the CPU oracle fingerprints the same clip, the two hash sets are
  compared (Jaccard) and the ORACLE's hashes are looked up in the same index — recognised ids / offsets / counts
  must be identical.
Prints one JSON object; run on a B200: python tools/config3_accuracy.py > gpurun_out/config3.json
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
FS = 44100


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tracks", type=int, default=2714)
    ap.add_argument("--seconds", type=int, default=180)
    ap.add_argument("--fan", type=int, default=15)
    ap.add_argument("--oracle-sample", type=int, default=24)
    ap.add_argument("--topn", type=int, default=3)
    ap.add_argument("--latency-clips", type=int, default=20, help="single-clip latency section (0 = skip)")
    args = ap.parse_args()

    import torch
    from bench import synth_tracks_gpu
    from oracle import sia_oracle as O
    from shazam_b200 import _native as N
    from shazam_b200.database import GPUDatabase
    from shazam_b200.fingerprinter import Fingerprinter

    dev = torch.device("cuda", 0)
    L = args.seconds * FS
    stride = (L + 7) // 8 * 8
    fp = Fingerprinter(0, max_chunk_frames=262144)
    p = fp.params(Fs=FS, fan_value=args.fan, amp_min=10)
    frames = N.num_frames(L)
    db = GPUDatabase(device=0, capacity_rows=int(args.tracks * frames * 5.5 * (args.fan - 1)) + (1 << 20))

    # ---- ingest --------------------------------------------------------------------------------------
    t_ingest = time.perf_counter()
    batch = 256
    pcm = torch.empty(batch * stride, dtype=torch.int16, device=dev)
    rows = pcm.view(batch, stride)
    keep_tracks = {}                     # a few tracks kept on the host for the oracle sample
    total_rows = 0
    for b0 in range(0, args.tracks, batch):
        nb = min(batch, args.tracks - b0)
        synth_tracks_gpu(dev, 5_000_000 + b0, nb, L, [rows[i, :L] for i in range(nb)])
        res = fp.fingerprint_device(pcm, np.arange(nb, dtype=np.int64) * stride, np.full(nb, L, np.int64), p)
        for i in range(nb):
            s, e = int(res.starts[i]), int(res.starts[i + 1])
            sid = db.insert_song(f"track{b0 + i:05d}", f"{b0 + i:040X}", e - s)
            db.insert_hashes_array(sid, res.hash[s:e], res.t1[s:e])
            db.set_song_fingerprinted(sid)
            total_rows += e - s
        del res
    stored = db.get_num_fingerprints()
    torch.cuda.synchronize()
    t_ingest = time.perf_counter() - t_ingest

    # ---- queries ----------------------------------------------------------------------------------------
    rng = np.random.default_rng(2020)
    out = {"config": {"tracks": args.tracks, "track_seconds": args.seconds, "fan_value": args.fan,
                      "index_rows": stored, "rows_inserted": total_rows, "topn": args.topn},
           "ingest_seconds": round(t_ingest, 2), "ingest_audio_s_per_s": args.tracks * args.seconds / t_ingest,
           "runs": []}
    sample_ids = rng.choice(args.tracks, min(args.oracle_sample, args.tracks), replace=False)
    for clip_s in (5, 15):
        for snr in (None, 10.0, 0.0):
            n = clip_s * FS
            cstride = (n + 7) // 8 * 8
            starts_sec = rng.integers(0, args.seconds - clip_s, args.tracks)
            clips = torch.zeros(args.tracks * cstride, dtype=torch.int16, device=dev)
            cv = clips.view(args.tracks, cstride)
            g = torch.Generator(device=dev)
            g.manual_seed(99 + clip_s + int(snr or -1))
            for b0 in range(0, args.tracks, batch):
                nb = min(batch, args.tracks - b0)
                synth_tracks_gpu(dev, 5_000_000 + b0, nb, L, [rows[i, :L] for i in range(nb)])
                for i in range(nb):
                    s0 = int(starts_sec[b0 + i]) * FS
                    sig = rows[i, s0:s0 + n].double()
                    if snr is not None:
                        # band-limited noise: white noise through a 64-tap Hann FIR
                        w = torch.randn(n + 63, device=dev, generator=g, dtype=torch.float64)
                        k = torch.hann_window(64, periodic=False, device=dev, dtype=torch.float64)
                        noise = torch.nn.functional.conv1d(w[None, None], k[None, None])[0, 0]
                        rms_s = torch.sqrt(torch.mean(sig ** 2))
                        rms_n = torch.sqrt(rms_s ** 2 / (10 ** (snr / 10)))
                        sig = sig + noise * (rms_n / torch.sqrt(torch.mean(noise ** 2)))
                    cv[b0 + i, :n] = torch.clamp(torch.round(sig), -32768, 32767).to(torch.int16)
            cstarts = np.arange(args.tracks, dtype=np.int64) * cstride
            clens = np.full(args.tracks, n, np.int64)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            q = fp.fingerprint_device(clips, cstarts, clens, p)
            torch.cuda.synchronize()
            t_fp = time.perf_counter() - t0
            t0 = time.perf_counter()
            song, diff, cnt, rws, nres = db.index.query_batch(q.hash, q.t1, q.starts, args.topn)
            torch.cuda.synchronize()
            t_q = time.perf_counter() - t0
            song = song.cpu().numpy(); diff = diff.cpu().numpy(); cnt = cnt.cpu().numpy(); nres = nres.cpu().numpy()
            truth = np.arange(args.tracks) + 1
            exp_off = np.round(starts_sec * FS / 2048.0).astype(np.int64)
            ok_song = (nres > 0) & (song[:, 0] == truth)
            ok_off = ok_song & (np.abs(diff[:, 0] - exp_off) <= 1)
            # oracle agreement on a sample
            jac, same, same_ids = [], 0, 0
            for tid in sample_ids:
                clip = cv[tid, :n].cpu().numpy()
                oh, ot = O.fingerprint_arrays(clip, FS, args.fan)
                gh = q.hash[int(q.starts[tid]):int(q.starts[tid + 1])].cpu().numpy()
                gt = q.t1[int(q.starts[tid]):int(q.starts[tid + 1])].cpu().numpy()
                sa = set(zip(map(bytes, gh), gt.tolist())); sb = set(zip(map(bytes, oh), ot.tolist()))
                jac.append(len(sa & sb) / max(1, len(sa | sb)))
                o = db.index.query_batch(torch.from_numpy(oh).to(dev), torch.from_numpy(ot).to(dev),
                                         np.array([0, len(ot)], np.int64), args.topn)
                k = int(o[4][0])
                a = (o[0][0, :k].tolist(), o[1][0, :k].tolist(), o[2][0, :k].tolist())
                kk = int(nres[tid])
                b = (song[tid, :kk].tolist(), diff[tid, :kk].tolist(), cnt[tid, :kk].tolist())
                same += a == b
                same_ids += a[:2] == b[:2]          # recognised song ids and offsets (north_star's criterion)
            out["runs"].append({
                "clip_seconds": clip_s, "snr_db": snr, "queries": args.tracks,
                "accuracy_song": float(ok_song.mean()), "accuracy_song_and_offset": float(ok_off.mean()),
                "mean_hashes_per_query": float(np.diff(q.starts).mean()),
                "fingerprint_ms_per_query": 1e3 * t_fp / args.tracks, "query_align_ms_per_query": 1e3 * t_q / args.tracks,
                "oracle_sample": len(sample_ids), "hash_set_jaccard_min": float(min(jac)),
                "hash_set_jaccard_mean": float(np.mean(jac)), "identical_ids_and_offsets_vs_oracle_hashes": same_ids,
                "identical_ids_offsets_counts_vs_oracle_hashes": same})
            del clips, q
    # ---- single-clip latency: the reference's main flow (recognizer.py:355-398) -----------------------------------
    # one 5 s stereo recording -> fingerprint() per channel -> find_matches -> align_matches, through the drop-in names
    # (shazam_b200.compat / shazam_b200.recognize: Python lists of (hex20, offset) tuples and the SELECT statement the
    # reference builds), and through the array fast path (Fingerprinter -> recognize_batch).  The reference's own
    # recorded run: 0.347 s per 2-channel clip (tests_csv; 0.28 s fingerprint + query + align on a 13 M-row table).
    if args.latency_clips > 0:
        from shazam_b200 import compat, recognize
        compat.set_fingerprinter(fp)
        recognize.set_database(db)
        n = 5 * FS
        synth_tracks_gpu(dev, 5_000_000, min(batch, args.tracks), L, [rows[i, :L] for i in range(min(batch, args.tracks))])
        lat_ref, lat_fast, ok = [], [], 0
        for k in range(args.latency_clips):
            tid = k % min(batch, args.tracks)
            s0 = (7 + 3 * k) % (args.seconds - 5) * FS
            left = rows[tid, s0:s0 + n].cpu().numpy()
            right = np.clip(left.astype(np.int32) + rng.integers(-40, 40, n), -32768, 32767).astype(np.int16)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res_ref = []
            for channel in (left, right):                          # recognizer.py:369-398
                hashes = set(compat.fingerprint(channel, Fs=FS, fan_value=args.fan))
                matches, dedup, _ = recognize.find_matches(hashes)
                res_ref.append(recognize.align_matches(matches, dedup, len(hashes), args.topn))
            lat_ref.append(time.perf_counter() - t0)
            t0 = time.perf_counter()
            b = fp.fingerprint_tracks([left, right], Fs=FS, fan_value=args.fan)
            res_fast = recognize.recognize_batch([b.track(0), b.track(1)], args.topn)
            lat_fast.append(time.perf_counter() - t0)
            strip = lambda rs: [(r["song_id"], r["offset"], r["hashes_matched_in_input"]) for r in rs]
            ok += int(all(strip(a) == strip(c) for a, c in zip(res_ref, res_fast)) and res_ref[0][0]["song_id"] == tid + 1)
        out["single_clip_latency"] = {
            "clips": args.latency_clips, "channels": 2, "clip_seconds": 5, "index_rows": stored,
            "drop_in_names_ms_median": 1e3 * float(np.median(lat_ref)), "drop_in_names_ms_max": 1e3 * float(np.max(lat_ref)),
            "array_fast_path_ms_median": 1e3 * float(np.median(lat_fast)), "array_fast_path_ms_max": 1e3 * float(np.max(lat_fast)),
            "identical_results_and_correct_song": ok,
            "reference_recorded_seconds": 0.347,
            "what": "fingerprint x2 channels + find_matches + align_matches per recording (recognizer.py:355-398); drop-in "
                    "names = compat.fingerprint / recognize.find_matches / recognize.align_matches with Python tuple lists; "
                    "fast path = Fingerprinter.fingerprint_tracks + recognize_batch"}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
