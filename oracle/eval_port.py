"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the translation-error windows of the reference evaluator
(mobileposer/evaluate.py:66-92).  Pinned to the live reference by oracle/make_golden_eval.py ->
tests/golden/tran_windows.npz (tests/test_evaluate.py).  Never imported by the product path.

For one sequence with ground-truth root translation `tran_t` [T, 3] and predicted `tran_p` [T, 3]:
  * `move[j+1] = move[j] + |tran_t[j+1] - tran_t[j]|`, accumulated sequentially in fp32             (evaluate.py:68-71)
  * for each window w = 1..7 m a two-pointer sweep collects the (start, end) pairs over which the ground truth moves
    at least w metres -- for every distinct `end` only the first `start` that reaches it            (evaluate.py:73-83)
  * per pair `|dt - dp| / (move[end] - move[start]) * w`, and the mean over the pairs (none -> the window is skipped
    for this sequence)                                                                              (evaluate.py:85-92)
"""
from __future__ import annotations

import numpy as np

WINDOWS = (1, 2, 3, 4, 5, 6, 7)


def move_distance(tran_t: np.ndarray) -> np.ndarray:
    t = np.asarray(tran_t, np.float32)
    d = t[1:] - t[:-1]
    v = np.sqrt((d * d).sum(axis=1, dtype=np.float32)).astype(np.float32)
    move = np.zeros(t.shape[0], np.float32)
    for j in range(v.shape[0]):
        move[j + 1] = np.float32(move[j] + v[j])
    return move


def frame_pairs(move: np.ndarray, window: int):
    pairs, start, end, n = [], 0, 1, move.shape[0]
    while end < n:
        if np.float32(move[end] - move[start]) < window:
            end += 1
        else:
            if not pairs or pairs[-1][1] != end:
                pairs.append((start, end))
            start += 1
    return pairs


def tran_window_errors(tran_p: np.ndarray, tran_t: np.ndarray, windows=WINDOWS):
    """-> (mean error per window [len(windows)] float32, NaN where the sequence has no pair; pair counts)."""
    tran_p, tran_t = np.asarray(tran_p, np.float32), np.asarray(tran_t, np.float32)
    move = move_distance(tran_t)
    out = np.full(len(windows), np.nan, np.float32)
    cnt = np.zeros(len(windows), np.int32)
    for k, w in enumerate(windows):
        pairs = frame_pairs(move, w)
        if not pairs:
            continue
        tot = np.float32(0)
        for s, e in pairs:
            d = (tran_t[e] - tran_t[s]) - (tran_p[e] - tran_p[s])
            err = np.float32(np.sqrt(np.float32((d * d).sum(dtype=np.float32)))) / np.float32(move[e] - move[s]) * np.float32(w)
            tot = np.float32(tot + np.float32(err))
        out[k] = tot / np.float32(len(pairs))
        cnt[k] = len(pairs)
    return out, cnt
