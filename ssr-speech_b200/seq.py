"""Host-side sequence surgery of `SSR_Speech.inference` (integer-only; numpy).

Mirrors, per utterance:
  * prologue  — reference models/ssr.py:604-626: span bookkeeping, `rearrange` (:381-406),
    delay pattern `shift`/`get_pattern_sequence` (:408-436,466-470), `insert_mask` (:472-494),
    `cat_y` (:496-502) and the truncation right before the first generation slot (:622-626);
  * epilogue  — reference models/ssr.py:774-812: `revert_pattern_sequence` (:438-464), EOG column
    drop, splice with the kept context, `marks` / `masks` construction.

The reference does this with per-element tensor writes in Python double loops; here it is
vectorised numpy on [K, T] int64 arrays.  The decode loop in between runs on the GPU.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np

from .config import SSRConfig


@dataclass
class PreparedUtterance:
    prompt_tokens: np.ndarray                 # [K, Y0-1] int64: audio tokens fed before the first <mts>
    mask_intervals: List[Tuple[int, int]]
    non_mask_intervals: List[Tuple[int, int]]
    num_spans: int
    y: np.ndarray                             # [K, T] original codes (after aug_context concat)
    out_len: int = 0                          # aug_context offset (ssr.py:579-593)


def delay_pattern(tokens: np.ndarray, special: int) -> np.ndarray:
    """[K, n] -> [K, n+K-1]; codebook q delayed by q, holes = `special` (ssr.py:408-436)."""
    K, n = tokens.shape
    out = np.full((K, n + K - 1), special, dtype=np.int64)
    for q in range(K):
        out[q, q:q + n] = tokens[q]
    return out


def revert_delay_pattern(pattern: np.ndarray, special: int) -> np.ndarray:
    """[K, S] -> [K, S-K+1] (ssr.py:438-464)."""
    K, S = pattern.shape
    T = S - (K - 1)
    out = np.full((K, max(T, 0)), special, dtype=np.int64)
    for q in range(K):
        out[q, :] = pattern[q, q:q + T]
    return out


def prepare(cfg: SSRConfig, y: np.ndarray, mask_interval: Sequence[Sequence[int]],
            out_len: int = 0) -> PreparedUtterance:
    """y: [K, T] int64; mask_interval: M x 2 (frame units).  ssr.py:604-626."""
    K, T = y.shape
    assert K == cfg.n_codebooks, y.shape
    mi = [(int(a) + out_len, int(b) + out_len) for a, b in mask_interval]
    assert 1 <= len(mi) <= cfg.max_n_spans, f"1..{cfg.max_n_spans} spans supported, got {len(mi)}"
    starts = [a for a, _ in mi] + [T]
    ends = [0] + [b for _, b in mi]
    non_mask = list(zip(ends, starts))
    col = lambda v: np.full((K, 1), v, dtype=np.int64)
    # rearrange (ssr.py:381-406)
    segs = []
    for i, (s, e) in enumerate(non_mask):
        if i == 0:
            segs.append(col(cfg.sos) if s == e else np.concatenate([col(cfg.sos), y[:, s:e]], axis=1))
        elif i == len(non_mask) - 1:
            segs.append(col(cfg.eos) if s == e else np.concatenate([y[:, s:e], col(cfg.eos)], axis=1))
        else:
            segs.append(y[:, s:e])
    # masked segments (ground truth + eog) are built by the reference and then truncated away; only
    # their count matters at inference.
    shifted = [delay_pattern(s, cfg.empty_token) for s in segs]
    # insert_mask (ssr.py:472-494): kept_0 <m0> kept_1 <m1> ... kept_M | <m0> masked_0 ...
    parts = []
    for j, s in enumerate(shifted):
        parts.append(s)
        if j < len(shifted) - 1:
            parts.append(col(cfg.mts + j))
    prompt = np.concatenate(parts, axis=1)          # everything before the second <mts_0>
    return PreparedUtterance(prompt_tokens=prompt, mask_intervals=mi, non_mask_intervals=non_mask,
                             num_spans=len(mi), y=y, out_len=out_len)


def finalize(cfg: SSRConfig, prep: PreparedUtterance, spans: List[np.ndarray]):
    """spans[i]: [n_i, K] int64 tokens sampled for span i (every iteration incl. the EOG tail).
    Returns (res [K, T_new], marks [T_new], masks, non_mask_intervals) — ssr.py:774-812."""
    y = prep.y
    K = cfg.n_codebooks
    flat = []
    for sp in spans:
        pat = np.asarray(sp, dtype=np.int64).T                      # [K, n]
        assert pat.shape[0] == K, pat.shape
        un = revert_delay_pattern(pat, cfg.empty_token)
        assert un.shape[1] == pat.shape[1] - K + 1
        flat.append(un[:, :-1])                                     # remove eog column
    res, marks, masks = [], [], []
    tmp = 0
    nm = prep.non_mask_intervals
    for (s, e), gen in zip(nm, flat):
        res.append(y[:, s:e])
        masks.append((tmp, tmp + e - s))
        marks += [0] * (e - s)
        res.append(gen)
        tmp += e - s + gen.shape[1]
        marks += [1] * gen.shape[1]
    if y.shape[1] != nm[-1][1] + 1:                                 # ssr.py:797 (as written in the reference)
        s, e = nm[-1]
        res.append(y[:, s:e])
        masks.append((tmp, tmp + e - s))
        marks += [0] * (e - s)
    res = np.concatenate(res, axis=1)
    marks = np.asarray(marks, dtype=np.int64)
    non_mask = list(nm)
    if prep.out_len:
        o = prep.out_len
        res = res[:, o:]
        marks = marks[o:]
        masks = [(a - o, b - o) for a, b in masks]
        non_mask = [(a - o, b - o) for a, b in non_mask]
    return res, marks, masks, non_mask


def expected_steps(cfg: SSRConfig, x_len: int, prompt_len: int) -> int:
    """Upper bound on decode iterations for one span under the reference's length guard
    (`y_input.shape[1] > 10*x_len`, ssr.py:739): EOG is forced at the first iteration whose audio
    length exceeds 10*x_len, then K-1 more iterations drain the delay pattern."""
    y0 = prompt_len + 1                       # + <mts>
    j_star = max(10 * x_len - y0 + 1, 0) + 1  # iteration index (1-based) at which the guard fires
    return j_star + cfg.n_codebooks - 1


def loss_flags(cfg: SSRConfig, y: np.ndarray, predict_mask_token: bool, predict_all: bool) -> np.ndarray:
    """Loss masks of the training forward (reference models/ssr.py:330-345) for ONE utterance's dataset-prepared tokens y [K, T]
    (cut to its length).  The targets are y[:, 1:]; returns uint8 flags [K, T-1]: bit 0 = the position enters the cross entropy
    and the top-10 accuracy (`tmp_masks`), bit 1 = it counts as a token (`masks`).
      masks     = target is neither audio_pad nor empty (and below `mts` unless predict_mask_token)
      tmp_masks = masks, minus everything BEFORE each occurrence of the first mask-token id `mts` unless predict_all"""
    tg = np.asarray(y)[:, 1:]
    mask = (tg != cfg.audio_pad_token) & (tg != cfg.empty_token)
    if not predict_mask_token:
        mask &= tg < cfg.mts
    tmp = mask.copy()
    if not predict_all:
        is_mts = tg == cfg.mts
        # clearing [:t] for every occurrence t == clearing everything before the LAST occurrence in the row
        has = is_mts.any(axis=1)
        last = tg.shape[1] - 1 - np.argmax(is_mts[:, ::-1], axis=1)
        cols = np.arange(tg.shape[1])[None, :]
        tmp &= ~(has[:, None] & (cols < last[:, None]))
    return (tmp.astype(np.uint8) | (mask.astype(np.uint8) << 1))

