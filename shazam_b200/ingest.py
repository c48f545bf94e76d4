"""Ingest driver mirror (``__init__.py:70-113, 248-415``): walk a directory, decode, fingerprint
every channel, insert song + hashes with the reference's two-phase commit flag.

The reference fans files out over ``multiprocessing.Pool`` and fingerprints one channel per call;
here all channels of a batch of files go to the GPU as ONE batch (``Fingerprinter``), which is
the same data-parallel unit (whole tracks, independent).  Decode stays on the host: the
reference uses pydub/ffmpeg (absent here), this mirror reads PCM WAV with the standard library;
other containers must be decoded by the caller and fed to ``compat.fingerprint`` /
``Fingerprinter.fingerprint_tracks`` channel by channel.
"""
from __future__ import annotations

import fnmatch
import os
import wave
from hashlib import sha1
from time import time
from typing import Optional

import numpy as np

from . import compat
from .fingerprinter import FingerprintBatch

db = None                      # module global, like the reference (set_database)
FIELD_FILE_SHA1 = 2            # __init__.py:40: get_songs() rows are tuples, sha1 at index 2
DEFAULT_FAN_VALUE = compat.DEFAULT_FAN_VALUE
DEFAULT_AMP_MIN = compat.DEFAULT_AMP_MIN


def set_database(database) -> None:
    global db
    db = database


def unique_hash(file_path: str, block_size: int = 2 ** 20) -> str:
    """``__init__.py:305-323``: upper-hex SHA-1 of the file bytes."""
    s = sha1()
    with open(file_path, "rb") as f:
        while True:
            buf = f.read(block_size)
            if not buf:
                break
            s.update(buf)
    return s.hexdigest().upper()


def find_files(path: str, extensions):
    """``__init__.py:286-303``."""
    extensions = [e.replace(".", "") for e in extensions]
    results = []
    for dirpath, dirnames, files in os.walk(path):
        for extension in extensions:
            for f in fnmatch.filter(files, f"*.{extension}"):
                results.append((os.path.join(dirpath, f), extension))
    return results


def read(file_name: str, limit: Optional[int] = None):
    """``__init__.py:70-113`` for 16-bit PCM WAV: ``(channels, frame_rate, file_sha1)`` with the
    channels de-interleaved (``data[chn::n_channels]``) and optionally cut to ``limit`` seconds."""
    with wave.open(file_name, "rb") as w:
        if w.getsampwidth() != 2:
            raise ValueError(f"{file_name}: only 16-bit PCM WAV is decoded here")
        nch, rate = w.getnchannels(), w.getframerate()
        nframes = w.getnframes() if not limit else min(w.getnframes(), int(limit) * rate)
        data = np.frombuffer(w.readframes(nframes), np.int16)
    channels = [np.ascontiguousarray(data[chn::nch]) for chn in range(nch)]
    return channels, rate, unique_hash(file_name)


def union_channels(batch: FingerprintBatch, first: int, count: int):
    """Set union of the (hash, offset) pairs of ``count`` consecutive tracks of a batch — the
    ``fingerprints |= set(hashes)`` of ``get_file_fingerprints`` (``__init__.py:256-265``).
    Returns (uint8[n,10], int32[n]) without duplicates."""
    s, e = int(batch.starts[first]), int(batch.starts[first + count])
    h = np.asarray(batch.hash[s:e])
    t = np.asarray(batch.t1[s:e])
    if count == 1 or len(t) == 0:
        return h, t                      # one channel never repeats a (hash, offset) pair
    rec = np.empty(len(t), dtype=[("h", "V10"), ("t", "<i4")])
    rec["h"] = np.frombuffer(np.ascontiguousarray(h).tobytes(), dtype="V10")
    rec["t"] = t
    _, keep = np.unique(rec, return_index=True)
    keep.sort()
    return h[keep], t[keep]


def union_channels_device(digests, offsets, scratch_index=None):
    """The same set union on the GPU: rows of all channels go through the index's own
    sort + adjacent-duplicate drop (UNIQUE(song_id, offset, hash) with one song id).  CUDA tensors in,
    (uint8[n,10], int32[n]) CUDA tensors out, in (hash, offset) order."""
    import torch
    from .database import FingerprintIndex
    n = int(offsets.numel())
    ix = scratch_index or FingerprintIndex(digests.device.index or 0, max(n, 1024))
    try:
        ix.insert(0, digests, offsets)
        ix.finalize()
        d, _, o = ix.export()
        return d, o
    finally:
        if scratch_index is None:
            ix.close()


def get_file_fingerprints_device(file_name: str, limit: Optional[int] = None, fan_value: int = DEFAULT_FAN_VALUE,
                                 amp_min=DEFAULT_AMP_MIN):
    """``get_file_fingerprints`` with everything after the decode on the GPU: the interleaved
    samples of a 16-bit PCM WAV go to the device once, are split per channel there, fingerprinted as one batch
    and set-unioned there.  Returns ``(digests uint8[n,10], offsets int32[n], file_sha1)`` as CUDA tensors in
    (hash, offset) order — the array form of the reference's ``set[(hex20, offset)]``."""
    with wave.open(file_name, "rb") as w:
        if w.getsampwidth() != 2:
            raise ValueError(f"{file_name}: only 16-bit PCM WAV is decoded here")
        nch, rate = w.getnchannels(), w.getframerate()
        nframes = w.getnframes() if not limit else min(w.getnframes(), int(limit) * rate)
        data = np.frombuffer(w.readframes(nframes), np.int16)
    fp = compat.get_fingerprinter()
    batch = fp.fingerprint_interleaved(data, nch, Fs=rate, fan_value=fan_value, amp_min=amp_min,
                                       connectivity=compat.CONNECTIVITY_MASK, nbhd=compat.PEAK_NEIGHBORHOOD_SIZE)
    d, o = union_channels_device(batch.hash, batch.t1)
    return d, o, unique_hash(file_name)


def get_file_fingerprints(file_name: str, limit: Optional[int] = None, print_output: bool = False):
    """``__init__.py:248-268``: ``(set[(hex20, offset)], file_sha1)``."""
    channels, fs, file_hash = read(file_name, limit)
    fingerprints = set()
    for channel in channels:
        fingerprints |= set(compat.fingerprint(channel, Fs=fs))
    return fingerprints, file_hash


def fingerprint_directory(path: str, extensions, nprocesses: Optional[int] = None, songhashes_set=None,
                          limit: Optional[int] = None, files_per_batch: int = 64,
                          fan_value: int = DEFAULT_FAN_VALUE, amp_min=DEFAULT_AMP_MIN):
    """``__init__.py:325-405``.  ``nprocesses`` is accepted for signature parity and ignored: the
    GPU batch is the pool.  Returns the number of songs inserted."""
    songhashes_set = set() if songhashes_set is None else songhashes_set
    todo = []
    for filename, _ in find_files(path, extensions):
        if unique_hash(filename) in songhashes_set:      # don't refingerprint (resume)
            print(f"{filename} already fingerprinted, continuing...")
            continue
        todo.append(filename)
    fp = compat.get_fingerprinter()
    start_t = time()
    done = 0
    for b0 in range(0, len(todo), files_per_batch):
        files = todo[b0:b0 + files_per_batch]
        by_rate = {}                                       # Fs enters the dB scale: one GPU batch per rate
        for f in files:
            try:
                channels, fs, file_hash = read(f, limit)
            except Exception as ex:                        # the reference prints and carries on
                print("Failed fingerprinting", f, ex)
                continue
            decoded, tracks = by_rate.setdefault(fs, ([], []))
            decoded.append((f, file_hash, len(tracks), len(channels)))
            tracks.extend(channels)
        for fs, (decoded, tracks) in by_rate.items():
            batch = fp.fingerprint_tracks(tracks, Fs=fs, fan_value=fan_value, amp_min=amp_min,
                                          connectivity=compat.CONNECTIVITY_MASK, nbhd=compat.PEAK_NEIGHBORHOOD_SIZE)
            for f, file_hash, first, count in decoded:
                song_name = os.path.splitext(os.path.basename(f))[0]
                h, t = union_channels(batch, first, count)
                sid = db.insert_song(song_name, file_hash, len(t))     # fingerprinted = 0
                db.insert_hashes_array(sid, h, t)
                db.set_song_fingerprinted(sid)                         # commit marker, __init__.py:381-386
                done += 1
        songhashes_set = load_fingerprinted_audio_hashes(songhashes_set)
    print("Total time to load {} songs: {}".format(done, time() - start_t))
    return done


def load_fingerprinted_audio_hashes(songhashes_set):
    """``__init__.py:407-415``."""
    for song in db.get_songs():
        songhashes_set.add(song[FIELD_FILE_SHA1])
    return songhashes_set
