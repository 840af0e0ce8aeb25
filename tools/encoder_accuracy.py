#!/usr/bin/env python
"""Encoder accuracy against the CPU oracle (fp32): latent error, RVQ index mismatches and the relative distance gap of every mismatch
(a mismatch is a near-tie when the oracle's distances to the two candidates differ by less than the latent noise).
    python tools/encoder_accuracy.py            # 3 x TF32 tensor-core encoder (default with precision="bf16")
    SSRB_ENC_FP32=1 python tools/encoder_accuracy.py   # fp32 CUDA-core encoder"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    from codec_oracle import CodecOracle
    from ssr_speech_b200.codec import WMEncodecModel
    from ssr_speech_b200.config import CodecConfig
    from ssr_speech_b200.synth import make_codec_state_dict
    g = np.load(os.path.join(ROOT, "tests", "golden", "codec_small.npz"))
    cfg = CodecConfig()
    sd = make_codec_state_dict(cfg, seed=int(g["weights_seed"]), codebook_mu=g["codebook_mu"], codebook_sigma=g["codebook_sigma"])
    o = CodecOracle(cfg, sd)
    m = WMEncodecModel(cfg, max_batch_chunk=32, precision="bf16")
    m.load_state_dict(sd)
    m.to("cuda")
    B, T = 6, 160000
    wav = torch.stack([0.1 * torch.randn(1, T, generator=torch.Generator().manual_seed(1234 + i)) for i in range(B)])
    codes, _, emb = m.encode(wav.cuda())
    worst_gap, n_bad, n_tot, worst_e = 0.0, 0, 0, 0.0
    for i in range(B):
        oc, _, oe = o.encode(wav[i:i + 1])
        e = float((emb[i:i + 1].cpu() - oe).abs().max() / oe.abs().max())
        worst_e = max(worst_e, e)
        got, want = codes[i:i + 1].cpu().numpy(), oc.numpy()
        res = oe.permute(0, 2, 1).reshape(-1, cfg.dimension).double()
        for q in range(cfg.n_q):
            E = o.codebook(q).double()
            dist = res.pow(2).sum(1, keepdim=True) - 2 * res @ E.t() + E.pow(2).sum(1)[None]
            for t in np.argwhere(got[0, q] != want[0, q]).ravel():
                d = dist[t]
                gap = float(abs(d[got[0, q, t]] - d[want[0, q, t]]) / d[want[0, q, t]].abs())
                worst_gap = max(worst_gap, gap)
                n_bad += 1
                print(f"utt {i} stage {q} frame {t}: rel gap {gap:.3e}")
            res = res - E[torch.from_numpy(want[0, q])]
        n_tot += want.size
    print(f"latent max-abs err / max|ref| = {worst_e:.3e}; index mismatches {n_bad} / {n_tot}; largest relative distance gap {worst_gap:.3e}")


if __name__ == "__main__":
    main()
