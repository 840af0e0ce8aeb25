"""TEST INFRASTRUCTURE — CPU oracle for the WM-Encodec path that brackets the decode loop.

Restates (torch CPU tensor arithmetic, fp32 or fp64) with the reference's state_dict as input:

  * weight-norm folding  w = g * v / ||v||  (norm over all dims but 0; for ConvTranspose1d dim 0 is the
    INPUT channel)                               torch legacy weight_norm; audiocraft/modules/conv.py:21-30
  * StreamableConv1d: asymmetric zero padding  left = total - total//2, right = total//2 + extra
                                                 audiocraft/modules/conv.py:47-53,185-201
  * StreamableConvTranspose1d: full transposed conv then trim left = total - total//2, right = total//2
                                                 audiocraft/modules/conv.py:221-243
  * SEANetResnetBlock (ELU, k3 conv, ELU, k1 conv, + identity skip)     modules/seanet.py:16-60
  * SEANetEncoder / SEANetDecoder (n_residual_layers = 1)               modules/seanet.py:63-258
  * StreamableLSTM: 2-layer LSTM (gate order i,f,g,o) + skip            modules/lstm.py:10-25
  * RVQ encode (argmax of -(|x|^2 - 2 x.E + |E|^2), first index) / decode (sum of rows)
                                                 quantization/core_vq.py:164-193,382-400; vq.py:87-103
  * WMSEANetDecoder.forward (skip encoder, wm_embed with max_norm=1 renorm, wm_proj*, decoder slices,
    wm_encoder + wm_predictor)                                          modules/seanet.py:555-600
  * WMEncodecModel.encode / decode / wmdecode (renormalize=False)       models/wmencodec.py:324-375

PARITY PIN: checked against the unmodified reference by `oracle/gen_golden.py` (the reference's own
tests only pin shapes: audiocraft/tests/modules/test_conv.py:151-203, test_seanet.py:18-36).
Only tests/, smoke() and bench.py's cpu_baseline / --impl reference legs may import this file.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F


def fold_weight_norm(g: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    n = v.flatten(1).norm(dim=1).view(-1, *([1] * (v.ndim - 1)))
    return v * (g / n)


def conv_out_len(T: int, k: int, s: int) -> int:
    """Output length of StreamableConv1d (dilation 1): ceil(T / s)  (test_conv.py:151-173)."""
    return math.ceil(T / s)


class CodecOracle:
    def __init__(self, cfg, state_dict: Dict[str, torch.Tensor], dtype=torch.float32):
        self.cfg = cfg
        self.dtype = dtype
        self.sd = {k: v.detach().to(dtype) if v.is_floating_point() else v for k, v in state_dict.items()}
        self._folded: Dict[str, torch.Tensor] = {}

    # -- primitives --------------------------------------------------------------------------
    def _w(self, prefix: str) -> torch.Tensor:
        if prefix not in self._folded:
            if prefix + "weight_g" in self.sd:
                self._folded[prefix] = fold_weight_norm(self.sd[prefix + "weight_g"], self.sd[prefix + "weight_v"])
            else:
                self._folded[prefix] = self.sd[prefix + "weight"]
        return self._folded[prefix]

    def conv(self, x: torch.Tensor, prefix: str, stride: int = 1) -> torch.Tensor:
        w = self._w(prefix)
        k = w.shape[-1]
        T = x.shape[-1]
        total = k - stride
        n_frames = (T - k + total) / stride + 1
        extra = (math.ceil(n_frames) - 1) * stride + (k - total) - T
        right = total // 2
        left = total - right
        x = F.pad(x, (left, right + extra))
        return F.conv1d(x, w, self.sd[prefix + "bias"], stride=stride)

    def convtr(self, x: torch.Tensor, prefix: str, stride: int) -> torch.Tensor:
        w = self._w(prefix)
        k = w.shape[-1]
        total = k - stride
        y = F.conv_transpose1d(x, w, self.sd[prefix + "bias"], stride=stride)
        right = total // 2
        left = total - right
        return y[..., left:y.shape[-1] - right]

    def resblock(self, x: torch.Tensor, prefix: str) -> torch.Tensor:
        h = self.conv(F.elu(x), prefix + "block.1.conv.conv.")
        h = self.conv(F.elu(h), prefix + "block.3.conv.conv.")
        return x + h

    def lstm(self, x: torch.Tensor, prefix: str) -> torch.Tensor:
        """x: [B, C, T]; 2 stacked layers, zero initial state, + skip."""
        B, C, T = x.shape
        inp = x.permute(2, 0, 1)                                   # [T, B, C]
        seq = inp
        for l in range(self.cfg.lstm):
            wih, whh = self.sd[f"{prefix}lstm.weight_ih_l{l}"], self.sd[f"{prefix}lstm.weight_hh_l{l}"]
            b = self.sd[f"{prefix}lstm.bias_ih_l{l}"] + self.sd[f"{prefix}lstm.bias_hh_l{l}"]
            pre = F.linear(seq, wih, b)                             # [T, B, 4C]
            h = torch.zeros(B, C, dtype=x.dtype)
            c = torch.zeros(B, C, dtype=x.dtype)
            outs = []
            for t in range(T):
                gates = pre[t] + F.linear(h, whh)
                i, f, g, o = gates.split(C, dim=-1)
                c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
                h = torch.sigmoid(o) * torch.tanh(c)
                outs.append(h)
            seq = torch.stack(outs, 0)
        return (seq + inp).permute(1, 2, 0)

    # -- encoder / decoder as lists of stages ------------------------------------------------------
    def encoder_stages(self, prefix: str):
        """Returns the encoder as a list of callables matching `model[i]` granularity groups
        [0:2], [2:5], [5:8], [8:11], [11:] used by WMSEANetDecoder.forward (seanet.py:559-574)."""
        r = list(reversed(self.cfg.ratios))
        p = prefix + "model."

        def s0(z):
            return self.resblock(self.conv(z, p + "0.conv.conv."), p + "1.")

        def mk(i, stride, res_idx):
            def f(z):
                z = self.conv(F.elu(z), p + f"{i}.conv.conv.", stride)
                return self.resblock(z, p + f"{res_idx}.")
            return f

        def last(z):
            z = self.conv(F.elu(z), p + "12.conv.conv.", r[3])
            z = self.lstm(z, p + "13.")
            return self.conv(F.elu(z), p + "15.conv.conv.")

        return [s0, mk(3, r[0], 4), mk(6, r[1], 7), mk(9, r[2], 10), last]

    def encoder(self, x: torch.Tensor, prefix: str = "encoder.") -> torch.Tensor:
        for st in self.encoder_stages(prefix):
            x = st(x)
        return x

    def decoder_stages(self, prefix: str):
        """Slices model[:4], [4:7], [7:10], [10:] of seanet.py:577-591."""
        r = list(self.cfg.ratios)
        p = prefix + "model."

        def d0(z):
            z = self.conv(z, p + "0.conv.conv.")
            z = self.lstm(z, p + "1.")
            return self.convtr(F.elu(z), p + "3.convtr.convtr.", r[0])

        def mk(res_idx, tr_idx, stride):
            def f(z):
                z = self.resblock(z, p + f"{res_idx}.")
                return self.convtr(F.elu(z), p + f"{tr_idx}.convtr.convtr.", stride)
            return f

        def last(z):
            z = self.resblock(z, p + "10.")
            z = self.convtr(F.elu(z), p + "12.convtr.convtr.", r[3])
            z = self.resblock(z, p + "13.")
            return self.conv(F.elu(z), p + "15.conv.conv.")

        return [d0, mk(4, 6, r[1]), mk(7, 9, r[2]), last]

    def decoder(self, z: torch.Tensor, prefix: str = "decoder.") -> torch.Tensor:
        for st in self.decoder_stages(prefix):
            z = st(z)
        return z

    # -- RVQ ---------------------------------------------------------------------------------------
    def codebook(self, q: int) -> torch.Tensor:
        return self.sd[f"quantizer.vq.layers.{q}._codebook.embed"]

    def rvq_encode(self, emb: torch.Tensor) -> torch.Tensor:
        """emb [B, 128, T] -> codes [B, n_q, T] int64."""
        B, D, T = emb.shape
        res = emb.permute(0, 2, 1).reshape(B * T, D)
        out = []
        for q in range(self.cfg.n_q):
            E = self.codebook(q)
            dist = -(res.pow(2).sum(1, keepdim=True) - 2 * res @ E.t() + E.t().pow(2).sum(0, keepdim=True))
            ind = dist.max(dim=-1).indices
            res = res - E[ind]
            out.append(ind.view(B, T))
        return torch.stack(out, 1)

    def rvq_decode(self, codes: torch.Tensor) -> torch.Tensor:
        """codes [B, n_q, T] -> [B, 128, T]."""
        z = None
        for q in range(codes.shape[1]):
            e = self.codebook(q)[codes[:, q]]
            z = e if z is None else z + e
        return z.permute(0, 2, 1)

    # -- model API ---------------------------------------------------------------------------------
    @torch.no_grad()
    def encode(self, wav: torch.Tensor):
        emb = self.encoder(wav.to(self.dtype))
        return self.rvq_encode(emb), None, emb

    @torch.no_grad()
    def decode(self, codes: torch.Tensor) -> torch.Tensor:
        return self.decoder(self.rvq_decode(codes))

    def _wm_embed(self, labels: torch.Tensor) -> torch.Tensor:
        w = self.sd["wmdecoder.wm_embed.weight"].clone()           # Embedding(max_norm=True -> 1.0)
        n = w.norm(dim=1, keepdim=True)
        w = torch.where(n > 1.0, w * (1.0 / (n + 1e-7)), w)
        return w[labels].transpose(2, 1)                            # [B, e, T]

    @torch.no_grad()
    def wmdecode(self, codes: torch.Tensor, marks: torch.Tensor, wav: torch.Tensor):
        """Returns (wav_out [B,1,320T], mark_logits [B,T,2])."""
        r = list(self.cfg.ratios)
        x = self.rvq_decode(codes)
        z = wav.to(self.dtype)
        enc = self.encoder_stages("wmdecoder.skip_encoder.")
        z = enc[0](z)
        skips, labels = [], []
        reps = [r[0] * r[1] * r[2], r[0] * r[1], r[0], 1]
        for st, rep in zip(enc[1:], reps):
            z = st(z)
            skips.append(z)
            labels.append(torch.repeat_interleave(marks, rep, dim=-1))
        dec = self.decoder_stages("wmdecoder.")
        for i, st in enumerate(dec):
            cat = torch.cat([skips.pop(), self._wm_embed(labels.pop())], dim=1)
            out = self.conv(F.elu(cat), f"wmdecoder.wm_proj{i}.1.conv.conv.") + x
            x = st(out)
        m = self.encoder(x, "wmdecoder.wm_encoder.")
        m = self.conv(F.elu(m), "wmdecoder.wm_predictor.1.conv.conv.")
        return x, m.transpose(2, 1)
