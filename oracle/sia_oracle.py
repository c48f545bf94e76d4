"""CPU oracle for SIA's fingerprint-and-match path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.  Nothing under
``shazam_b200/`` imports it: the product path is CUDA-only and fails loudly when
the extension is missing.

What this restates (reference = CarlosArturoMe/shazam, paths relative to it):

* ``spectrogram_db``  — ``mlab.specgram(..)[0]`` + the dB transform as called at
  ``__init__.py:232-241``.  matplotlib is an UN-VENDORED, UNPINNED third-party
  dependency (``environment.yml:8``) and is not installed in this image, so this
  part is a restatement of matplotlib's published ``_spectral_helper``
  (mode 'psd', detrend none, ``scale_by_freq=True``, one-sided) — SURVEY.md
  Appendix A.  Pin: ``tests/test_oracle.py`` cross-checks it against
  ``scipy.signal.spectrogram`` with the equivalent settings.  No reference test
  holds a golden spectrogram, so for THIS stage parity is "pinned to scipy, not
  to a reference vector".
* ``get_2D_peaks``    — ``__init__.py:116-177`` (scipy.ndimage calls restated
  one for one) plus ``peaks_bruteforce``, an independent clipped-window
  statement used to validate the simplification the CUDA kernel relies on.
* ``generate_hashes`` — ``__init__.py:179-210``.
* ``fingerprint``     — ``__init__.py:212-245``.
* ``FingerprintTable``/``return_matches``/``align_matches`` —
  ``mysql_database.py:32-59,167-233`` (schema + set semantics of
  ``INSERT IGNORE`` under ``UNIQUE(song_id, offset, hash)``),
  ``recognizer.py:222-271`` and ``recognizer.py:289-338``.

Pin for peaks / hashes / match: ``tests/golden/make_golden.py`` executes the
reference's OWN function bodies (AST-extracted from ``/root/reference``) and
commits their outputs under ``tests/golden/``; ``tests/test_oracle.py`` checks
this oracle against those vectors bit for bit.
"""
from __future__ import annotations

import hashlib
from itertools import groupby
from operator import itemgetter

import numpy as np

# constants: __init__.py:40-51, recognizer.py:38,68
RATE = 44100
DEFAULT_FS = 44100
DEFAULT_WINDOW_SIZE = 4096
DEFAULT_OVERLAP_RATIO = 0.5
DEFAULT_FAN_VALUE = 5
DEFAULT_AMP_MIN = 10
CONNECTIVITY_MASK = 2
PEAK_NEIGHBORHOOD_SIZE = 10
MIN_HASH_TIME_DELTA = 0
MAX_HASH_TIME_DELTA = 200
FINGERPRINT_REDUCTION = 20
TOPN = 2


# --------------------------------------------------------------------------- #
# a-2 / a-3 : spectrogram in dB
# --------------------------------------------------------------------------- #
def num_frames(n_samples: int, nfft: int = 4096, noverlap: int = 2048) -> int:
    """Frames mlab.specgram produces: short inputs are zero-padded to one frame,
    the trailing partial frame is dropped."""
    n = max(int(n_samples), nfft)
    return (n - noverlap) // (nfft - noverlap)


def specgram_psd(x, Fs: float = RATE, nfft: int = 4096, noverlap: int = 2048) -> np.ndarray:
    """One-sided PSD, float64 ``[nfft//2+1][T]`` — ``mlab.specgram(...)[0]`` as
    called at ``__init__.py:232-237`` (window_hanning, detrend none)."""
    x = np.asarray(x)
    if x.ndim != 1:
        raise ValueError("channel_samples must be 1-D")
    if len(x) < nfft:
        x = np.concatenate([x, np.zeros(nfft - len(x), dtype=x.dtype)])
    win = np.hanning(nfft)  # symmetric Hann, denominator nfft-1
    step = nfft - noverlap
    frames = np.lib.stride_tricks.sliding_window_view(x, nfft)[::step]  # (T, nfft)
    X = np.fft.fft(frames * win[None, :], n=nfft, axis=1)[:, : nfft // 2 + 1]
    P = (np.conj(X) * X).real
    P[:, 1:-1] *= 2.0  # every bin except DC and Nyquist (nfft is even)
    P /= Fs
    P /= (win ** 2).sum()
    return np.ascontiguousarray(P.T)


def spectrogram_db(x, Fs: float = RATE, nfft: int = 4096, noverlap: int = 2048) -> np.ndarray:
    """``10*log10(P)`` with ``P == 0 -> 0.0`` (``__init__.py:241``)."""
    P = specgram_psd(x, Fs, nfft, noverlap)
    return 10 * np.log10(P, out=np.zeros_like(P), where=(P != 0))


# --------------------------------------------------------------------------- #
# a-4 : peaks
# --------------------------------------------------------------------------- #
def neighborhood(connectivity: int = CONNECTIVITY_MASK, size: int = PEAK_NEIGHBORHOOD_SIZE) -> np.ndarray:
    from scipy.ndimage import generate_binary_structure, iterate_structure
    return iterate_structure(generate_binary_structure(2, connectivity), size)


def get_2D_peaks(arr2D: np.ndarray, amp_min=DEFAULT_AMP_MIN,
                 connectivity: int = CONNECTIVITY_MASK, size: int = PEAK_NEIGHBORHOOD_SIZE):
    """``__init__.py:116-177``: max-filter equality, XOR with eroded zero
    background, ``> amp_min``; returns ``[(f, t), ...]`` in np.where order
    (freq-major)."""
    from scipy.ndimage import binary_erosion, maximum_filter
    nb = neighborhood(connectivity, size)
    local_max = maximum_filter(arr2D, footprint=nb) == arr2D
    eroded_background = binary_erosion(arr2D == 0, structure=nb, border_value=1)
    detected = local_max != eroded_background
    amps = arr2D[detected]
    freqs, times = np.where(detected)
    keep = np.where(amps.flatten() > amp_min)
    return list(zip(freqs[keep], times[keep]))


def peaks_bruteforce(arr2D: np.ndarray, amp_min=DEFAULT_AMP_MIN,
                     connectivity: int = CONNECTIVITY_MASK, size: int = PEAK_NEIGHBORHOOD_SIZE):
    """Independent statement of what the CUDA kernel computes: window clipped to
    the array, ``v == max(window) and not all(window == 0) and v > amp_min``.
    O(F*T*441) — small inputs only."""
    F, T = arr2D.shape
    out = []
    for f in range(F):
        for t in range(T):
            v = arr2D[f, t]
            if not v > amp_min:
                continue
            ok, allzero = True, True
            for df in range(-size, size + 1):
                ff = f + df
                if ff < 0 or ff >= F:
                    continue
                w = size if connectivity == 2 else size - abs(df)
                lo, hi = max(0, t - w), min(T, t + w + 1)
                seg = arr2D[ff, lo:hi]
                if seg.max() > v:
                    ok = False
                    break
                if allzero and np.any(seg != 0):
                    allzero = False
            if ok and not allzero:
                out.append((f, t))
    return out


# --------------------------------------------------------------------------- #
# a-5 : hashes
# --------------------------------------------------------------------------- #
def generate_hashes(peaks, fan_value: int = DEFAULT_FAN_VALUE):
    """``__init__.py:179-210``: stable sort by time, ``fan_value-1`` partners,
    ``0 <= dt <= 200``, sha1("f1|f2|dt")[:20] paired with t1."""
    peaks = sorted(peaks, key=itemgetter(1))  # stable, like list.sort(key=...)
    out = []
    n = len(peaks)
    for i in range(n):
        f1, t1 = peaks[i]
        for j in range(1, fan_value):
            if i + j < n:
                f2, t2 = peaks[i + j]
                dt = t2 - t1
                if MIN_HASH_TIME_DELTA <= dt <= MAX_HASH_TIME_DELTA:
                    h = hashlib.sha1(f"{f1}|{f2}|{dt}".encode("utf-8"))
                    out.append((h.hexdigest()[0:FINGERPRINT_REDUCTION], t1))
    return out


def fingerprint(channel_samples, Fs: int = RATE, wsize: int = DEFAULT_WINDOW_SIZE,
                wratio: float = DEFAULT_OVERLAP_RATIO, fan_value: int = DEFAULT_FAN_VALUE,
                amp_min=DEFAULT_AMP_MIN, connectivity: int = CONNECTIVITY_MASK):
    """``__init__.py:212-245``."""
    arr2D = spectrogram_db(channel_samples, Fs, wsize, int(wsize * wratio))
    return generate_hashes(get_2D_peaks(arr2D, amp_min, connectivity), fan_value)


def fingerprint_arrays(channel_samples, Fs=RATE, fan_value=DEFAULT_FAN_VALUE, amp_min=DEFAULT_AMP_MIN,
                       connectivity=CONNECTIVITY_MASK):
    """``fingerprint`` with array outputs: ``uint8[N,10]`` digests, ``int32[N]`` t1."""
    hs = fingerprint(channel_samples, Fs, fan_value=fan_value, amp_min=amp_min, connectivity=connectivity)
    if not hs:
        return np.zeros((0, 10), np.uint8), np.zeros((0,), np.int32)
    h = np.frombuffer(bytes.fromhex("".join(x[0] for x in hs)), np.uint8).reshape(-1, 10).copy()
    t = np.array([int(x[1]) for x in hs], np.int32)
    return h, t


# --------------------------------------------------------------------------- #
# a-8 .. a-11 : table, lookup, vote
# --------------------------------------------------------------------------- #
class FingerprintTable:
    """In-memory stand-in for the MySQL ``songs``/``fingerprints`` tables
    (``mysql_database.py:32-59``).  ``insert_hashes`` has INSERT IGNORE set
    semantics on (song_id, offset, hash); ``select_multiple`` returns every stored
    ``(HEXUPPER, song_id, offset)`` whose hash is in the IN-list
    (``recognizer.py:60-64``)."""

    def __init__(self):
        self.songs = {}          # song_id -> dict
        self.rows = {}           # HEXUPPER -> list[(song_id, offset)]
        self._unique = set()     # (song_id, offset, HEXUPPER)
        self._next_id = 1

    def insert_song(self, song_name, file_hash, total_hashes):
        sid = self._next_id
        self._next_id += 1
        self.songs[sid] = {"song_name": song_name, "file_sha1": file_hash,
                           "total_hashes": total_hashes, "fingerprinted": 0}
        return sid

    def insert_hashes(self, song_id, hashes, batch_size=1000):
        for hsh, offset in hashes:
            key = (song_id, int(offset), hsh.upper())
            if key in self._unique:
                continue
            self._unique.add(key)
            self.rows.setdefault(hsh.upper(), []).append((song_id, int(offset)))

    def set_song_fingerprinted(self, song_id):
        self.songs[song_id]["fingerprinted"] = 1

    def get_song_by_id(self, song_id):
        s = self.songs[song_id]
        return {"song_name": s["song_name"], "total_hashes": s["total_hashes"], "file_sha1": s["file_sha1"]}

    def select_multiple(self, hex_upper_list):
        for h in hex_upper_list:
            for sid, off in self.rows.get(h, ()):
                yield h, sid, off

    def num_rows(self):
        return len(self._unique)


class ArrayFingerprintTable(FingerprintTable):
    """The same stand-in for tables too large for a dict of tuples: the ``fingerprints`` rows live in sorted numpy
    arrays (the DB engine's index); ``select_multiple`` still hands the rows to the caller ONE BY ONE as
    ``(HEXUPPER, song_id, offset)`` — the cursor iteration of ``recognizer.py:259`` whose Python loop is what the
    reference spends its query time in."""

    def __init__(self, digests, song_ids, offsets):
        super().__init__()
        d = np.ascontiguousarray(digests, np.uint8).reshape(-1, 10)
        hi = d[:, :8].copy().view(">u8").reshape(-1).astype(np.uint64)
        lo = d[:, 8:10].copy().view(">u2").reshape(-1).astype(np.uint64)
        rec = np.empty(len(hi), dtype=[("hi", "<u8"), ("lo", "<u8"), ("s", "<i8"), ("o", "<i8")])
        rec["hi"], rec["lo"], rec["s"], rec["o"] = hi, lo, np.asarray(song_ids, np.int64), np.asarray(offsets, np.int64)
        rec = np.unique(rec)                      # UNIQUE(song_id, offset, hash) + (hash, song, offset) order
        self._hi, self._lo = rec["hi"].copy(), rec["lo"].copy()
        self._s, self._o = rec["s"].copy(), rec["o"].copy()

    @classmethod
    def from_sorted(cls, hi, lo, song_ids, offsets):
        """Rows that are already unique and in (hash, song_id, offset) order (hi: first 8 digest bytes as big-endian
        uint64, lo: last 2 as an integer) — e.g. an export of the GPU index — without the sort of the constructor."""
        t = cls.__new__(cls)
        FingerprintTable.__init__(t)
        t._hi = np.ascontiguousarray(hi, np.uint64)
        t._lo = np.ascontiguousarray(lo, np.uint64)
        t._s = np.ascontiguousarray(song_ids, np.int64)
        t._o = np.ascontiguousarray(offsets, np.int64)
        return t

    def insert_hashes(self, song_id, hashes, batch_size=1000):
        raise NotImplementedError("ArrayFingerprintTable is built from arrays")

    def select_multiple(self, hex_upper_list):
        for h in hex_upper_list:
            v = int(h, 16)
            hi, lo = np.uint64(v >> 16), np.uint64(v & 0xffff)
            a = int(np.searchsorted(self._hi, hi, "left"))
            b = int(np.searchsorted(self._hi, hi, "right"))
            if b > a:
                lo_run = self._lo[a:b]
                a, b = a + int(np.searchsorted(lo_run, lo, "left")), a + int(np.searchsorted(lo_run, lo, "right"))
            for sid, off in zip(self._s[a:b].tolist(), self._o[a:b].tolist()):
                yield h, sid, off

    def num_rows(self):
        return len(self._s)


def return_matches(table: FingerprintTable, hashes, batch_size: int = 1000):
    """``recognizer.py:222-271``."""
    mapper = {}
    for hsh, offset in hashes:
        mapper.setdefault(hsh.upper(), []).append(offset)
    values = list(mapper.keys())
    dedup_hashes = {}
    results = []
    for index in range(0, len(values), batch_size):
        for hsh, sid, offset in table.select_multiple(values[index: index + batch_size]):
            dedup_hashes[sid] = dedup_hashes.get(sid, 0) + 1
            for song_sampled_offset in mapper[hsh]:
                results.append((sid, offset - song_sampled_offset))
    return results, dedup_hashes


def best_offsets(matches, topn: int = TOPN):
    """The vote inside ``align_matches`` (``recognizer.py:303-310``):
    ``[(song_id, offset_diff, count), ...]`` best first; per song the first
    (smallest-diff) maximum; equal counts stay in ascending song_id order."""
    sorted_matches = sorted(matches, key=lambda m: (m[0], m[1]))
    counts = [(*key, len(list(group))) for key, group in groupby(sorted_matches, key=lambda m: (m[0], m[1]))]
    songs_matches = sorted(
        [max(list(group), key=lambda g: g[2]) for key, group in groupby(counts, key=lambda c: c[0])],
        key=lambda c: c[2], reverse=True)
    return songs_matches[0:topn]


def align_matches(table: FingerprintTable, matches, dedup_hashes, queried_hashes, topn: int = TOPN):
    """``recognizer.py:289-338``: result dicts, same keys and roundings."""
    out = []
    for song_id, offset, _ in best_offsets(matches, topn):
        song = table.get_song_by_id(song_id)
        song_hashes = song.get("total_hashes", None)
        nseconds = round(float(offset) / DEFAULT_FS * DEFAULT_WINDOW_SIZE * DEFAULT_OVERLAP_RATIO, 5)
        hashes_matched = dedup_hashes[song_id]
        out.append({
            "song_id": song_id,
            "song_name": song.get("song_name", None).encode("utf8"),
            "input_total_hashes": queried_hashes,
            "fingerprinted_hashes_in_db": song_hashes,
            "hashes_matched_in_input": hashes_matched,
            "input_confidence": round(hashes_matched / queried_hashes, 2),
            "fingerprinted_confidence": round(hashes_matched / song_hashes, 2),
            "offset": offset,
            "offset_seconds": nseconds,
            "file_sha1": song.get("file_sha1", None).encode("utf8"),
        })
    return out


# --------------------------------------------------------------------------- #
# synthetic workload (SURVEY.md §8d Config 2) and the noise mixer
# --------------------------------------------------------------------------- #
def synth_track(seed: int, n_samples: int = 7_938_000, fs: int = RATE) -> np.ndarray:
    """Deterministic synthetic 'music': 6 voices, a note change every 11 025
    samples, pitch 110*2^(U{0..59}/12) Hz, amplitude U(500, 4000), harmonics
    (1, 1/2, 1/4), + N(0, 200^2) noise, clipped to int16."""
    rng = np.random.default_rng(seed)
    seg = 11025
    nseg = -(-n_samples // seg)
    t = np.arange(n_samples, dtype=np.float64) / fs
    x = np.zeros(n_samples, np.float64)
    for _ in range(6):
        semis = rng.integers(0, 60, nseg)
        amps = rng.uniform(500, 4000, nseg)
        f = np.repeat(110.0 * 2.0 ** (semis / 12.0), seg)[:n_samples]
        a = np.repeat(amps, seg)[:n_samples]
        ph = 2 * np.pi * f * t
        x += a * (np.sin(ph) + 0.5 * np.sin(2 * ph) + 0.25 * np.sin(3 * ph))
    x += rng.normal(0, 200, n_samples)
    return np.clip(np.rint(x), -32768, 32767).astype(np.int16)


def mix_noise(signal: np.ndarray, noise: np.ndarray, snr_db: float) -> np.ndarray:
    """``get_noise_from_sound`` (``recognizer_test.py:426-435``): scale ``noise``
    so RMS_s/RMS_n = 10^(SNR/20), add, return float64."""
    signal = np.asarray(signal, np.float64)
    noise = np.asarray(noise, np.float64)
    rms_s = np.sqrt(np.mean(signal ** 2))
    rms_n = np.sqrt(rms_s ** 2 / (10 ** (snr_db / 10)))
    rms_cur = np.sqrt(np.mean(noise ** 2))
    return signal + noise * (rms_n / rms_cur)
