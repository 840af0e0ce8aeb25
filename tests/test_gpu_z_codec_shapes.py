"""CUDA WM-Encodec kernels (fp32 parity mode, through the C ABI) against the oracle on the shapes of
tests/test_codec_oracle_vs_reference_sweep.py — a single frame, odd frame counts, batch 3, all watermark patterns, silence — where
the oracle itself is pinned against the unmodified reference in the build container.  Tolerances as in test_gpu_codec.py
(SURVEY §8c item 4)."""
import os

import numpy as np
import pytest
import torch

from codec_oracle import CodecOracle
from test_codec_oracle_vs_reference_sweep import CASES, signals
from test_gpu_codec import load, near_tie_only

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def small(gold_dir):
    return load(gold_dir, "codec_small.npz")


@pytest.mark.parametrize("B,frames,kind,mark_kind", CASES)
def test_codec_kernels_match_oracle(small, B, frames, kind, mark_kind):
    g, cfg, sd, m = small
    o = CodecOracle(cfg, sd)
    T = frames * 320
    wav = signals(B, T, kind, seed=frames * 7 + B)
    gen = torch.Generator().manual_seed(frames)
    marks = {"zeros": torch.zeros(B, frames, dtype=torch.long), "ones": torch.ones(B, frames, dtype=torch.long),
             "alt": (torch.arange(frames) % 2)[None].repeat(B, 1), "rand": torch.randint(0, 2, (B, frames), generator=gen),
             "block": torch.cat([torch.zeros(B, frames // 2, dtype=torch.long), torch.ones(B, frames - frames // 2, dtype=torch.long)], 1)}[mark_kind]
    ocodes, _, oemb = o.encode(wav)
    codes, scale, emb = m.encode(wav.cuda())
    assert scale is None and tuple(codes.shape) == (B, 4, frames)
    scale_e = max(float(oemb.abs().max()), 1e-6)
    assert float((emb.cpu() - oemb).abs().max()) <= 1e-4 * scale_e
    assert torch.equal(m.quantize(oemb.cuda()).cpu(), ocodes)                       # RVQ bit-exact given the oracle's latents
    assert near_tie_only(o, oemb.numpy(), codes.cpu().numpy(), ocodes.numpy())
    odec = o.decode(ocodes)
    dec = m.decode(ocodes.cuda())
    assert tuple(dec.shape) == tuple(odec.shape) == (B, 1, T)
    assert float((dec.cpu() - odec).abs().max()) <= 1e-4 * max(float(odec.abs().max()), 1e-6)
    owm, omlog = o.wmdecode(ocodes, marks, wav)
    wm, mlog = m.wmdecode(ocodes.cuda(), marks.cuda(), wav.cuda())
    assert float((wm.cpu() - owm).abs().max()) <= 1e-4 * max(float(owm.abs().max()), 1e-6)
    assert float((mlog.cpu() - omlog).abs().max()) <= 1e-4 * max(float(omlog.abs().max()), 1.0)
