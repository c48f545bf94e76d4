"""Noise-robustness harness: the accuracy / confusion-matrix report against scikit-learn (what the
reference calls, recognizer_test.py:501-503) on the CPU; the GPU SNR mixer against the oracle's mix_noise."""
import csv
import os

import numpy as np
import pytest

from oracle import sia_oracle as O


def test_classification_summary_equals_sklearn(tmp_path):
    sk = pytest.importorskip("sklearn.metrics")
    from shazam_b200 import evaluate
    rng = np.random.default_rng(4)
    names = [f"song {i:02d}" for i in range(12)]
    y_true = [names[i] for i in rng.integers(0, 12, 200)]
    y_pred = [t if rng.random() < 0.8 else names[int(rng.integers(0, 12))] for t in y_true]
    y_pred[5] = "never played"                                  # a label that only appears among the predictions
    s = evaluate.classification_summary(y_true, y_pred)
    assert np.array_equal(s["confusion_matrix"], sk.confusion_matrix(y_true, y_pred))
    assert s["accuracy"] == pytest.approx(sk.accuracy_score(y_true, y_pred))
    want = sk.classification_report(y_true, y_pred, output_dict=True, zero_division=0)
    assert set(want) == set(s["report"])
    for key, val in want.items():
        if isinstance(val, dict):
            for k2, v2 in val.items():
                assert s["report"][key][k2] == pytest.approx(v2), (key, k2)
        else:
            assert s["report"][key] == pytest.approx(val)
    # the CSV files of generate_csv_results (recognizer_test.py:437-513)
    played = [f"songs/album/{n}.mp3" for n in y_true]
    assert evaluate.track_name(played[0]) == y_true[0]
    times = [{"song_start_time": i, "fingerprint_times": 0.1, "query_time": 0.2, "align_time": 0.3, "total_time": 0.6}
             for i in range(len(played))]
    out = evaluate.generate_csv_results(played, y_pred, times, [[] for _ in played], record_seconds=5, snr=10,
                                        out_dir=str(tmp_path), stamp="01-01-2026_00-00-00")
    assert out["files"][0] == "shazam_results_01-01-2026_00-00-00_200records_5seconds_10SNR_atSong1.csv"
    rows = list(csv.DictReader(open(tmp_path / out["files"][0])))
    assert len(rows) == 200 and sum(int(r["correct"]) for r in rows) == out["correct"] == sum(a == b for a, b in zip(y_true, y_pred))
    assert all(os.path.exists(tmp_path / f) for f in out["files"])
    acc = list(csv.reader(open(tmp_path / out["files"][4])))
    assert float(acc[1][1]) == pytest.approx(out["accuracy"])


def test_report_files_equal_the_reference_script(tmp_path):
    """``generate_csv_results`` (recognizer_test.py:437-513) was executed from the reference's source by
    tests/golden/make_golden.py (csv_cases.json): the results CSV, the 0/1 crosstab (``CM_``, incl. the empty cells of
    a column added for a predicted name that is no played track), the sklearn confusion matrix (``CMSK_``), the
    classification report (``CRSK_``) and the accuracy score (``ASSK_``) — same file names, same bytes."""
    import json
    from shazam_b200 import evaluate
    cases = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "csv_cases.json")))
    assert len(cases) >= 3
    for case in cases:
        out_dir = tmp_path / str(case["case"])
        out_dir.mkdir()
        out = evaluate.generate_csv_results(case["songs_to_recognize"], case["recognized_song_names"], case["times"],
                                            case["final_results_arr"], case["record_seconds"],
                                            snr=case["snr"] if case["add_noise"] else None, iteration=case["iteration"],
                                            out_dir=str(out_dir), stamp=case["stamp"])
        assert sorted(out["files"]) == sorted(case["files"]), case["case"]
        for name, want in case["files"].items():
            got = open(out_dir / name, newline="").read()
            assert got.replace("\r\n", "\n") == want.replace("\r\n", "\n"), (case["case"], name, got, want)


@pytest.mark.gpu
def test_mix_noise_device_equals_oracle(native_lib):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from shazam_b200 import evaluate
    rng = np.random.default_rng(12)
    n, stride = 5 * 44100, 5 * 44100 + 4
    clips = np.zeros((6, stride), np.int16)
    noise = np.zeros((6, n + 100), np.float32)
    for c in range(6):
        clips[c, :n] = O.synth_track(60 + c, n) * (1 if c != 4 else 9)      # clip 4 saturates after mixing
        noise[c] = np.convolve(rng.normal(0, 1, n + 163), np.hanning(64), "valid").astype(np.float32)
    for snr in (10.0, 0.0, -3.0):
        got, scale = evaluate.mix_noise_device(torch.from_numpy(clips).cuda(), torch.from_numpy(noise).cuda(), n, snr,
                                               return_scale=True)
        got = got.cpu().numpy(); scale = scale.cpu().numpy()
        for c in range(6):
            mixed = O.mix_noise(clips[c, :n], noise[c, :n], snr)               # recognizer_test.py:426-435
            want = np.clip(np.rint(mixed), -32768, 32767).astype(np.int16)
            rms_s = np.sqrt(np.mean(clips[c, :n].astype(np.float64) ** 2))
            rms_n = np.sqrt(np.mean(noise[c, :n].astype(np.float64) ** 2))
            assert scale[c] == pytest.approx(rms_s / 10 ** (snr / 20) / rms_n, rel=1e-12)
            diff = np.abs(got[c, :n].astype(np.int32) - want.astype(np.int32))
            # identical up to the summation order of the two RMS values (a sample exactly on a rounding boundary)
            assert diff.max() <= 1 and (diff != 0).mean() < 1e-4, (snr, c, int(diff.max()), float((diff != 0).mean()))
            assert np.all(got[c, n:] == 0)
        assert np.abs(got[4].astype(np.int32)).max() == 32768 or got[4].max() == 32767        # saturated, not wrapped
