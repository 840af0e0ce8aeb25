"""TEST INFRASTRUCTURE — CPU oracle for the SSR-Speech autoregressive decode path.

A from-scratch restatement (torch CPU tensor arithmetic, fp32 or fp64) of what the reference
computes in `SSR_Speech.inference` (reference models/ssr.py:504-812) and the modules below it:

  * token embeddings + additive sinusoidal PE scaled by alpha
        models/modules/embedding.py:22-48, :51-98; models/ssr.py:191-198,596-600
  * pre-norm decoder layer  x += SA(LN1 x); x += FFN(LN2 x), final LN
        models/modules/transformer.py:58-75,321-343,386-388,473-488
  * packed-QKV multi-head attention, 1/sqrt(dh) scaling, plain causal mask over [text ; audio]
        models/modules/activation.py:83-89,536-637; models/ssr.py:227-257 (mask == triu)
  * 4 prediction heads Linear-GELU(erf)-Linear on the last position     models/ssr.py:175-179,687-689
  * CFG mix every cfg_stride-th step                                    models/ssr.py:690-696
  * logit rules / EOG bookkeeping / silence rule                        models/ssr.py:698-754
  * top-k / top-p filtering and sampling                                models/ssr.py:26-86

PARITY PIN: this oracle is checked against the *unmodified reference executed in the build
container* by `oracle/gen_golden.py`, which also writes the golden fixtures under `tests/golden/`
(the reference has no tests or golden vectors of its own for this path — SURVEY §4).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this file.  It is never on the product path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F


def sinusoid_table(n_pos: int, dim: int) -> torch.Tensor:
    """pe[p, 0::2] = sin(p*div), pe[p, 1::2] = cos(p*div) in fp32 (embedding.py:76-92)."""
    pos = torch.arange(0, n_pos, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, dim, 2, dtype=torch.float32) * -(math.log(10000.0) / dim))
    pe = torch.zeros(n_pos, dim)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


def filter_top_k_top_p(logits: torch.Tensor, top_k: int, top_p: float) -> torch.Tensor:
    """[K, V] -> filtered copy (ssr.py:26-68)."""
    logits = logits.clone()
    neg = -float("inf")
    if top_k > 0:
        k = min(max(top_k, 1), logits.shape[-1])
        kth = torch.topk(logits, k)[0][..., -1, None]
        logits[logits < kth] = neg
    if top_p < 1.0:
        srt, idx = torch.sort(logits, descending=True)
        cum = torch.cumsum(F.softmax(srt, dim=-1), dim=-1)
        rem = cum > top_p
        rem[..., 1:] = rem[..., :-1].clone()
        rem[..., 0] = False
        mask = torch.zeros_like(rem).scatter(1, idx, rem)
        logits[mask] = neg
    return logits


@dataclass
class StepTrace:
    raw_logits: torch.Tensor      # [R, K, V] logits straight out of the heads (before CFG / rules)
    final_logits: torch.Tensor    # [K, V] after CFG + rules (+ temperature), before top-k/top-p
    probs: torch.Tensor           # [K, V] post-filter softmax the sample is drawn from
    samples: torch.Tensor         # [K] tokens after the EOG overrides


class LMOracle:
    def __init__(self, cfg, state_dict: Dict[str, torch.Tensor], dtype=torch.float32,
                 round_weights_to_bf16: bool = False, round_acts_to_bf16: bool = False):
        """round_*_to_bf16: emulate the production kernels' storage precision (weights / GEMM operand
        activations / KV rounded to bf16, all arithmetic still fp32|fp64) to tighten bf16-mode tolerances."""
        self.cfg = cfg
        self.dtype = dtype
        self.round_acts = round_acts_to_bf16
        sd = {}
        for k, v in state_dict.items():
            v = v.detach().to(torch.float32)
            if round_weights_to_bf16 and v.ndim == 2 and "embedding" not in k:
                v = v.to(torch.bfloat16).to(torch.float32)
            sd[k] = v.to(dtype)
        self.sd = sd
        self.pe = sinusoid_table(4000, cfg.d_model)
        self.alpha_t = sd["text_positional_embedding.alpha"]
        self.alpha_a = sd["audio_positional_embedding.alpha"]

    # -- building blocks ---------------------------------------------------------------------
    def _ra(self, t):
        return t.to(torch.bfloat16).to(self.dtype) if self.round_acts else t

    def _pe(self, n):
        if n > self.pe.shape[0]:
            self.pe = sinusoid_table(n, self.cfg.d_model)
        return self.pe[:n].to(self.dtype)

    def embed_text(self, x: torch.Tensor) -> torch.Tensor:          # [Lx] -> [Lx, D]
        e = self.sd["text_embedding.word_embeddings.weight"][x]
        return e + self.alpha_t * self._pe(x.shape[0])

    def embed_audio_tokens(self, toks: torch.Tensor) -> torch.Tensor:  # [K, T] -> [T, D] (no PE)
        e = 0
        for k in range(self.cfg.n_codebooks):
            ek = self.sd[f"audio_embedding.{k}.word_embeddings.weight"][toks[k]]
            e = ek if k == 0 else e + ek
        return e

    def _ln(self, x, prefix):
        return F.layer_norm(x, (self.cfg.d_model,), self.sd[prefix + ".weight"], self.sd[prefix + ".bias"], 1e-5)

    def layer(self, n: int, x: torch.Tensor, kv: Optional[tuple]):
        """x: [q, D] new positions (appended after the cached ones).  Returns (x_out, (K, V)) with
        K, V: [H, S, dh] covering cached + new positions.  Causal within the new block."""
        cfg, sd = self.cfg, self.sd
        p = f"decoder.layers.{n}."
        H, dh = cfg.nhead, cfg.head_dim
        h = self._ra(self._ln(x, p + "norm1"))
        qkv = F.linear(h, sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"])
        q, k, v = qkv.split(cfg.d_model, dim=-1)
        q = q.view(-1, H, dh).transpose(0, 1)                      # [H, q, dh]
        k = self._ra(k).view(-1, H, dh).transpose(0, 1)
        v = self._ra(v).view(-1, H, dh).transpose(0, 1)
        if kv is not None:
            k = torch.cat([kv[0], k], dim=1)
            v = torch.cat([kv[1], v], dim=1)
        nq, S = q.shape[1], k.shape[1]
        att = torch.matmul(q, k.transpose(1, 2)) / math.sqrt(dh)    # [H, q, S]
        causal = torch.ones(nq, S, dtype=torch.bool).triu(S - nq + 1)
        att = att.masked_fill(causal, -float("inf"))
        o = torch.matmul(torch.softmax(att, dim=-1), v)             # [H, q, dh]
        o = self._ra(o.transpose(0, 1).reshape(nq, cfg.d_model))
        x = x + F.linear(o, sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"])
        h = self._ra(self._ln(x, p + "norm2"))
        h = self._ra(torch.relu(F.linear(h, sd[p + "linear1.weight"], sd[p + "linear1.bias"])))
        x = x + F.linear(h, sd[p + "linear2.weight"], sd[p + "linear2.bias"])
        return x, (k, v)

    def stack(self, x: torch.Tensor, cache: Optional[list]):
        new_cache = []
        for n in range(self.cfg.num_decoder_layers):
            x, kv = self.layer(n, x, None if cache is None else cache[n])
            new_cache.append(kv)
        return self._ln(x, "decoder.norm"), new_cache

    def heads(self, h: torch.Tensor) -> torch.Tensor:               # [D] -> [K, V]
        out = []
        hh = self._ra(h)
        for k in range(self.cfg.n_codebooks):
            p = f"predict_layer.{k}."
            z = F.gelu(F.linear(hh, self.sd[p + "0.weight"], self.sd[p + "0.bias"]))
            out.append(F.linear(self._ra(z), self.sd[p + "2.weight"], self.sd[p + "2.bias"]))
        return torch.stack(out, 0)

    # -- teacher-forced logits ------------------------------------------------------------------
    def teacher_forced_logits(self, x: torch.Tensor, audio_tokens: torch.Tensor) -> torch.Tensor:
        """x [Lx], audio_tokens [K, Ty]  ->  logits [Ty, K, V] for every audio position."""
        xi = self.embed_text(x)
        yi = self.embed_audio_tokens(audio_tokens) + self.alpha_a * self._pe(audio_tokens.shape[1])
        h, _ = self.stack(torch.cat([xi, yi], 0), None)
        h = h[x.shape[0]:]
        return torch.stack([self.heads(h[t]) for t in range(h.shape[0])], 0)

    # -- training forward / loss ------------------------------------------------------
    @staticmethod
    def loss_masks(cfg, y: torch.Tensor, predict_mask_token: bool = True, predict_all: bool = False):
        """models/ssr.py:333-347 on one utterance's y [K, T] (already cut to its length): the targets are y shifted by one,
        `mask` counts a target unless it is audio_pad / empty (and, without predict_mask_token, a mask token); `tmp_mask`
        additionally drops everything BEFORE each occurrence of the first mask-token id (`mts`) unless predict_all.
        Returns (targets [K, T-1], mask, tmp_mask)."""
        tg = y[:, 1:]
        mask = (tg != cfg.audio_pad_token) & (tg != cfg.empty_token)
        if not predict_mask_token:
            mask = mask & (tg < cfg.mts)
        tmp = mask.clone()
        if not predict_all:
            for k, t in (tg == cfg.mts).nonzero(as_tuple=False).tolist():
                tmp[k, :t] = False
        return tg, mask, tmp

    @torch.no_grad()
    def forward_loss(self, x: torch.Tensor, x_lens: torch.Tensor, y: torch.Tensor, y_lens: torch.Tensor,
                     predict_mask_token: bool = True, predict_all: bool = False, codebook_weight=None):
        """SSR_Speech.forward (models/ssr.py:280-379) in eval mode: x [B, S] int64, y [B, K, T] int64 (dataset-prepared: mask
        tokens, eog, audio_pad padding), lengths [B].  Padded positions are masked as keys in the reference, so every utterance
        is run at its own length.  Returns the reference's dict: loss = sum_k mean-CE_k * ntokens_k * weight_k, top10acc,
        top10acc_by_codebook, effective_ntoken."""
        cfg = self.cfg
        K = cfg.n_codebooks
        B = x.shape[0]
        nll = [[] for _ in range(K)]
        hit = [[] for _ in range(K)]
        ntok = [0] * K
        for b in range(B):
            xl, yl = int(x_lens[b]), int(y_lens[b])
            yb = y[b, :, :yl]
            logits = self.teacher_forced_logits(x[b, :xl], yb)[:-1].to(torch.float64)          # [T-1, K, V]
            tg, mask, tmp = self.loss_masks(cfg, yb, predict_mask_token, predict_all)
            for k in range(K):
                lk, tk = logits[:, k][tmp[k]], tg[k][tmp[k]]
                nll[k].append(F.cross_entropy(lk, tk, reduction="none"))
                hit[k].append((lk.topk(10, dim=-1).indices == tk[:, None]).any(-1))
                ntok[k] += int(mask[k].sum())
        cw = [1.0] * K if codebook_weight is None else list(codebook_weight)
        loss_k = [torch.cat(nll[k]).mean() for k in range(K)]
        acc_k = [torch.cat(hit[k]).double().mean() for k in range(K)]
        by_cb = [float(acc_k[k]) * ntok[k] for k in range(K)]
        return {"loss": float(sum(float(loss_k[k]) * ntok[k] * cw[k] for k in range(K))), "top10acc": float(sum(by_cb)),
                "top10acc_by_codebook": by_cb, "effective_ntoken": int(sum(ntok)),
                "loss_by_codebook": [float(v) for v in loss_k], "ntokens_by_codebook": ntok}

    # -- the decode loop ----------------------------------------------------------------------------
    @torch.no_grad()
    def inference(self, x: torch.Tensor, prompt_tokens: torch.Tensor, num_spans: int,
                  top_k: int = 0, top_p: float = 0.8, temperature: float = 1.0, stop_repetition: int = -1,
                  silence_tokens: Sequence[int] = (1388, 1898, 131), cfg_coef: float = 1.5, cfg_stride: int = 1,
                  aug_text: bool = False, uncond_x: Optional[torch.Tensor] = None,
                  noise: Optional[torch.Tensor] = None, trace: Optional[List[StepTrace]] = None,
                  max_steps: Optional[int] = None, incremental: bool = True):
        """One utterance.  x [Lx] int64; prompt_tokens [K, Y0-1] (seq.prepare().prompt_tokens).

        noise: optional [n_steps, K, V] Exp(1) variates; sample = argmax(p / noise) — exactly what
        torch.multinomial(num_samples=1) computes internally from its own exponential_() draw.  When
        None, torch.multinomial with the global CPU generator is used (as the reference does).
        Returns list (per span) of [n_i, K] int64 arrays of sampled tokens.
        """
        cfg = self.cfg
        K, V = cfg.n_codebooks, cfg.n_audio_tokens
        assert cfg_coef >= 1.0
        rows_x = [x]
        if aug_text:
            if uncond_x is None:   # ssr.py:574 — drawn on the global CPU generator
                uncond_x = torch.randint(0, cfg.n_text_tokens, (1, x.shape[0]))[0]
            rows_x.append(uncond_x)
        R = len(rows_x)
        x_in = [self.embed_text(r) for r in rows_x]
        Lx = x.shape[0]
        emb_y = self.embed_audio_tokens(prompt_tokens)              # [Y0-1, D], shared by cond/uncond rows
        caches = [None] * R
        n_cached = 0
        spans, step_no = [], 0
        for idx in range(num_spans):
            cur = []
            prev_token, consec, num_gen, num_eog, cfg_tag = None, 0, 0, 0, 1
            mts = torch.full((K, 1), cfg.mts + idx, dtype=torch.long)
            emb_y = torch.cat([emb_y, self.embed_audio_tokens(mts)], 0)
            while True:
                Ty = emb_y.shape[0]
                y_in = emb_y + self.alpha_a * self._pe(Ty)
                raw = []
                for r in range(R):
                    if incremental and caches[r] is not None:
                        h, caches[r] = self.stack(y_in[-1:], caches[r])
                    else:
                        h, c = self.stack(torch.cat([x_in[r], y_in], 0), None)
                        caches[r] = c if incremental else None
                    raw.append(self.heads(h[-1]))
                raw = torch.stack(raw, 0)                           # [R, K, V]
                if aug_text:
                    if cfg_tag == cfg_stride:
                        logits = cfg_coef * raw[0] + (1 - cfg_coef) * raw[1]
                        cfg_tag = 1
                    else:
                        cfg_tag += 1
                        logits = raw[0].clone()
                else:
                    logits = raw[0].clone()
                logits[:, cfg.eos] = -10000.0
                logits[:, cfg.sos] = -10000.0
                logits[:, cfg.mts:cfg.mts + cfg.max_n_spans] = -10000.0
                if num_gen < K - 1:
                    logits[num_gen + 1:, cfg.empty_token] = 10000.0
                if num_eog > 0:
                    logits[num_eog + 1:, cfg.eog] = -10000.0
                    logits[num_eog + 1:, cfg.empty_token] = -10000.0
                else:
                    logits[1:, cfg.eog] = -10000.0
                    if stop_repetition > 0 and prev_token in silence_tokens and consec > stop_repetition:
                        f = consec - (stop_repetition - 1)
                        if logits[0, prev_token] < 0:
                            logits[0, prev_token] = logits[0, prev_token] * f
                        else:
                            logits[0, prev_token] = logits[0, prev_token] / f
                lg = logits / temperature if temperature != 1.0 else logits
                filt = filter_top_k_top_p(lg, top_k, top_p)
                probs = F.softmax(filt, dim=-1)
                if noise is not None:
                    samples = torch.argmax(probs / noise[step_no].to(probs.dtype), dim=-1)
                else:
                    samples = torch.multinomial(probs.float() if probs.dtype != torch.float32 else probs, 1)[:, 0]
                samples = samples.clone()
                if num_eog > 0:
                    samples[:num_eog] = cfg.empty_token
                    samples[num_eog] = cfg.eog
                    num_eog += 1
                else:
                    # argmax over the filtered logits, like the reference: with temperature==1.0 its
                    # top_k_top_p_filtering mutates `logits` in place before the argmax at ssr.py:739
                    # (the top token always survives the filter, so this equals argmax(logits[0])).
                    if (samples[0] == cfg.eog or torch.argmax(filt[0]) == cfg.eog or Ty > Lx * 10):
                        samples[0] = cfg.eog
                        num_eog += 1
                    s0 = int(samples[0])
                    if s0 in silence_tokens and prev_token is not None and s0 == prev_token:
                        consec += 1
                    else:
                        consec = 0
                    prev_token = s0
                if trace is not None:
                    trace.append(StepTrace(raw.clone(), lg.clone(), probs.clone(), samples.clone()))
                num_gen += 1
                step_no += 1
                cur.append(samples.numpy().copy())
                if num_eog == K:
                    break
                if max_steps is not None and step_no >= max_steps:
                    spans.append(np.stack(cur, 0))
                    return spans
                emb_y = torch.cat([emb_y, self.embed_audio_tokens(samples.view(K, 1))], 0)
            spans.append(np.stack(cur, 0))
        return spans
