"""Noise-robustness harness: what the reference's experiment script does around the recogniser —
mix a noise recording into each query at a given SNR (``get_noise_from_sound``, ``recognizer_test.py:426-435, 554``)
and report per-run accuracy, confusion matrix and per-class precision / recall (``generate_csv_results``,
``recognizer_test.py:437-513``).  The mixing runs on the GPU for a whole batch of clips (``sia_mix_noise``); the
report is a few hundred labels of host arithmetic and writes the same CSV files the reference writes (its
loudspeaker / microphone loop and the mp3 decode stay outside).
"""
from __future__ import annotations

import csv
import ctypes as C
import datetime
import re
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _native as N


def mix_noise_device(clips: torch.Tensor, noise: torch.Tensor, n_samples: int, snr_db: float,
                     out: Optional[torch.Tensor] = None, return_scale: bool = False):
    """``signal + get_noise_from_sound(signal, noise, SNR)`` for a batch of clips on the GPU.

    ``clips``: int16 CUDA tensor ``[B, stride]`` (row c holds clip c in its first ``n_samples`` entries),
    ``noise``: float32 CUDA tensor ``[B, >= n_samples]`` — the noise segment cut for each clip
    (``recognizer_test.py:549-552``).  Returns int16 ``[B, stride]`` (saturated, rounded to nearest even)."""
    assert clips.is_cuda and clips.dtype == torch.int16 and clips.dim() == 2
    assert noise.is_cuda and noise.dtype == torch.float32 and noise.dim() == 2 and noise.shape[0] == clips.shape[0]
    clips = clips.contiguous(); noise = noise.contiguous()
    if out is None:
        out = torch.zeros_like(clips)
    scale = torch.empty(clips.shape[0], dtype=torch.float64, device=clips.device) if return_scale else None
    dev = clips.device.index or 0
    N.check(N.lib().sia_mix_noise(dev, C.c_void_p(clips.data_ptr()), clips.shape[1], C.c_void_p(noise.data_ptr()),
                                  noise.shape[1], clips.shape[0], int(n_samples), float(snr_db),
                                  C.c_void_p(out.data_ptr()), out.shape[1],
                                  C.c_void_p(scale.data_ptr()) if return_scale else None,
                                  C.c_void_p(torch.cuda.current_stream(clips.device).cuda_stream)))
    return (out, scale) if return_scale else out


def track_name(path: str) -> str:
    """The label ``generate_csv_results`` derives from a played file's path (``recognizer_test.py:445-451``): two
    leading ``dir/`` components and the ``.mp3`` suffix dropped."""
    name = re.sub(r"^.*?/", "", path)
    name = re.sub(r"^.*?/", "", name)
    return name.replace(".mp3", "")


def classification_summary(y_true: Sequence[str], y_pred: Sequence[str]) -> Dict:
    """Confusion matrix (rows = actual, columns = predicted, labels sorted — sklearn's convention, which the
    reference uses at ``recognizer_test.py:501-503``), accuracy score and the per-class precision / recall / f1 /
    support report with its macro and weighted averages."""
    y_true = [str(v) for v in y_true]
    y_pred = [str(v) for v in y_pred]
    assert len(y_true) == len(y_pred)
    labels = sorted(set(y_true) | set(y_pred))
    index = {l: i for i, l in enumerate(labels)}
    cm = np.zeros((len(labels), len(labels)), np.int64)
    for t, p in zip(y_true, y_pred):
        cm[index[t], index[p]] += 1
    tp = np.diag(cm).astype(np.float64)
    support = cm.sum(1).astype(np.float64)
    predicted = cm.sum(0).astype(np.float64)
    # the same arithmetic, operation for operation, as scikit-learn's precision_recall_fscore_support (the CSV files the
    # reference writes hold repr()s of these doubles): f1 from the confusion-matrix entries, 2 tp / (support + predicted)
    with np.errstate(divide="ignore", invalid="ignore"):
        precision = np.where(predicted > 0, tp / predicted, 0.0)
        recall = np.where(support > 0, tp / support, 0.0)
        f1 = np.where(support + predicted > 0, 2.0 * tp / (1.0 * support + predicted), 0.0)
    report = {l: {"precision": float(precision[i]), "recall": float(recall[i]), "f1-score": float(f1[i]),
                  "support": float(support[i])} for l, i in index.items()}
    n = float(len(y_true))
    accuracy = float(tp.sum() / n) if n else 0.0
    report["accuracy"] = accuracy
    report["macro avg"] = {"precision": float(precision.mean()), "recall": float(recall.mean()),
                           "f1-score": float(f1.mean()), "support": n}
    wavg = (lambda x: float(np.average(x, weights=support))) if n else (lambda x: 0.0)
    report["weighted avg"] = {"precision": wavg(precision), "recall": wavg(recall), "f1-score": wavg(f1), "support": n}
    return {"labels": labels, "confusion_matrix": cm, "accuracy": accuracy, "report": report}


def generate_csv_results(songs_to_recognize: Sequence[str], recognized_song_names: Sequence[str], times: Sequence[dict],
                         final_results_arr: Sequence, record_seconds: int, snr: Optional[float] = None,
                         iteration: int = 0, out_dir: str = ".", stamp: Optional[str] = None) -> Dict:
    """``recognizer_test.py:437-513``: one row per played file (correct = the recognised name equals the played
    track's name) in ``shazam_results_<stamp>_<n>records_<s>seconds[_<SNR>SNR]_atSong<k>.csv`` plus the ``CM_`` (the
    reference's 0/1 crosstab), ``CMSK_`` (confusion matrix), ``CRSK_`` (classification report) and ``ASSK_``
    (accuracy score) files next to it.  Returns the summary and the file names."""
    import os
    names = [track_name(s) for s in songs_to_recognize]
    rows = []
    for i, played in enumerate(songs_to_recognize):
        t = times[i]
        rows.append({"file_name_played": str(played), "file_name_result": str(recognized_song_names[i]),
                     "song_start_time": t.get("song_start_time"), "correct": int(names[i] == recognized_song_names[i]),
                     "fingerprint_times": t.get("fingerprint_times"), "query_time": t.get("query_time"),
                     "align_time": t.get("align_time"), "total_time": t.get("total_time"),
                     "final_results": final_results_arr[i]})
    stamp = stamp or datetime.datetime.now().strftime("%d-%m-%Y_%H-%M-%S")
    csv_name = f"shazam_results_{stamp}_{len(songs_to_recognize)}records_{record_seconds}seconds" + \
               (f"_{snr}SNR" if snr is not None else "") + f"_atSong{iteration + 1}.csv"
    columns = ["file_name_played", "file_name_result", "song_start_time", "correct", "fingerprint_times", "query_time",
               "align_time", "total_time", "final_results"]
    with open(os.path.join(out_dir, csv_name), "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=columns)
        w.writeheader()
        w.writerows(rows)
    summary = classification_summary(names, recognized_song_names)
    labels, cm = summary["labels"], summary["confusion_matrix"]
    # CM_: the reference's crosstab (recognizer_test.py:493-499) — crosstab(actual, actual), i.e. the diagonal cell of a
    # track is how often it was played; a miss zeroes it and puts a 1 in the predicted name's column.  A predicted name
    # that is no played track gets a NEW column, appended in order of appearance, whose untouched cells stay empty
    # (pandas fills an enlarged column with NaN and writes NaN as an empty field)
    actual = sorted(set(names))
    cols = list(actual)
    for t, p in zip(names, recognized_song_names):
        if t != p and p not in cols:
            cols.append(p)
    cross = {a: {c: (0 if c in actual else "") for c in cols} for a in actual}
    for a in names:
        cross[a][a] += 1
    for t, p in zip(names, recognized_song_names):
        if t != p:
            cross[t][t] = 0
            cross[t][p] = 1
    with open(os.path.join(out_dir, "CM_" + csv_name), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["Actual"] + cols)
        for a in actual:
            w.writerow([a] + [cross[a][c] for c in cols])
    with open(os.path.join(out_dir, "CMSK_" + csv_name), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([""] + list(range(len(labels))))
        for i, row in enumerate(cm.tolist()):
            w.writerow([i] + row)
    with open(os.path.join(out_dir, "CRSK_" + csv_name), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["", "precision", "recall", "f1-score", "support"])
        for key, val in summary["report"].items():
            if isinstance(val, dict):
                w.writerow([key, val["precision"], val["recall"], val["f1-score"], val["support"]])
            else:                                      # the accuracy row of sklearn's dict, transposed like the reference
                w.writerow([key, val, val, val, val])
    with open(os.path.join(out_dir, "ASSK_" + csv_name), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["", 0])
        w.writerow([0, summary["accuracy"]])
    summary["files"] = [csv_name, "CM_" + csv_name, "CMSK_" + csv_name, "CRSK_" + csv_name, "ASSK_" + csv_name]
    summary["correct"] = int(sum(r["correct"] for r in rows))
    return summary
