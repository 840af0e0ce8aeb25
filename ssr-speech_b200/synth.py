"""Seeded random-init checkpoints with the reference's state_dict layout.

No pretrained weights exist in the build container (reference `pretrained_models/test` is a
placeholder), so parity tests and the benchmark use these.  Every tensor is drawn from its own
CPU generator seeded by (seed, key), so any subset of keys can be regenerated independently and
the result does not depend on module construction order.  Magnitudes follow the reference's
default initialisers (xavier-uniform in_proj, activation.py:280-290; kaiming-uniform Linear/Conv;
N(0,1) embeddings) with small perturbations on LayerNorm/bias/weight_g so that every affine term
is exercised.
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict

import torch

from .config import CodecConfig, SSRConfig


def _gen(seed: int, key: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) & 0x7FFFFFFF)
    return g


def _uniform(seed, key, shape, bound):
    return (torch.rand(shape, generator=_gen(seed, key), dtype=torch.float32) * 2 - 1) * bound


def _normal(seed, key, shape, std=1.0, mean=0.0):
    return torch.randn(shape, generator=_gen(seed, key), dtype=torch.float32) * std + mean


def make_lm_state_dict(cfg: SSRConfig, seed: int = 0, pin_eog_bias: bool = False) -> "OrderedDict[str, torch.Tensor]":
    """state_dict for reference `SSR_Speech` (keys: SURVEY Appendix D).

    pin_eog_bias: benchmark checkpoints only — sets predict_layer[0][2].bias[eog] = -1e4 so that the
    generation length is decided by the reference's own length guard (ssr.py:739), SURVEY §8(d).
    """
    D, F, Hh, Va, Vt = cfg.d_model, cfg.ffn_dim, cfg.head_hidden, cfg.n_audio_tokens, cfg.n_text_tokens
    sd = OrderedDict()
    sd["text_embedding.word_embeddings.weight"] = _normal(seed, "text_emb", (Vt, D))
    for k in range(cfg.n_codebooks):
        sd[f"audio_embedding.{k}.word_embeddings.weight"] = _normal(seed, f"audio_emb{k}", (Va, D))
    sd["text_positional_embedding.alpha"] = torch.tensor([0.9])
    sd["audio_positional_embedding.alpha"] = torch.tensor([1.1])
    for n in range(cfg.num_decoder_layers):
        p = f"decoder.layers.{n}."
        sd[p + "self_attn.in_proj_weight"] = _uniform(seed, p + "qkv_w", (3 * D, D), math.sqrt(6.0 / (4 * D)))
        sd[p + "self_attn.in_proj_bias"] = _uniform(seed, p + "qkv_b", (3 * D,), 0.02)
        sd[p + "self_attn.out_proj.weight"] = _uniform(seed, p + "o_w", (D, D), 1.0 / math.sqrt(D))
        sd[p + "self_attn.out_proj.bias"] = _uniform(seed, p + "o_b", (D,), 0.02)
        sd[p + "linear1.weight"] = _uniform(seed, p + "w1", (F, D), 1.0 / math.sqrt(D))
        sd[p + "linear1.bias"] = _uniform(seed, p + "b1", (F,), 1.0 / math.sqrt(D))
        sd[p + "linear2.weight"] = _uniform(seed, p + "w2", (D, F), 1.0 / math.sqrt(F))
        sd[p + "linear2.bias"] = _uniform(seed, p + "b2", (D,), 1.0 / math.sqrt(F))
        for ln in ("norm1", "norm2"):
            sd[p + ln + ".weight"] = _normal(seed, p + ln + "w", (D,), 0.05, 1.0)
            sd[p + ln + ".bias"] = _normal(seed, p + ln + "b", (D,), 0.05, 0.0)
    sd["decoder.norm.weight"] = _normal(seed, "final_ln_w", (D,), 0.05, 1.0)
    sd["decoder.norm.bias"] = _normal(seed, "final_ln_b", (D,), 0.05, 0.0)
    for k in range(cfg.n_codebooks):
        p = f"predict_layer.{k}."
        sd[p + "0.weight"] = _uniform(seed, p + "0w", (Hh, D), 1.0 / math.sqrt(D))
        sd[p + "0.bias"] = _uniform(seed, p + "0b", (Hh,), 1.0 / math.sqrt(D))
        sd[p + "2.weight"] = _uniform(seed, p + "2w", (Va, Hh), 1.0 / math.sqrt(Hh))
        sd[p + "2.bias"] = _uniform(seed, p + "2b", (Va,), 1.0 / math.sqrt(Hh))
    if pin_eog_bias:
        sd["predict_layer.0.2.bias"][cfg.eog] = -1e4
    return sd


# ---------------------------------------------------------------------------------------------
# codec
# ---------------------------------------------------------------------------------------------

def _conv_entries(sd, seed, prefix, cout, cin, k, weight_norm=True, transpose=False):
    """NormConv1d / NormConvTranspose1d parameters (audiocraft/modules/conv.py:100-147).
    ConvTranspose1d weights are [in, out, k] and weight_norm's dim-0 is the *input* channel."""
    shape = (cin, cout, k) if transpose else (cout, cin, k)
    fan_in = (cout if transpose else cin) * k   # torch's fan_in for both layouts = shape[1]*k
    bound = 1.0 / math.sqrt(fan_in)
    v = _uniform(seed, prefix + "v", shape, bound)
    b = _uniform(seed, prefix + "b", (cout,), bound)
    if weight_norm:
        g = v.flatten(1).norm(dim=1).view(-1, 1, 1) * _normal(seed, prefix + "g", (shape[0], 1, 1), 0.05, 1.0)
        sd[prefix + "weight_g"] = g
        sd[prefix + "weight_v"] = v
    else:
        sd[prefix + "weight"] = v
    sd[prefix + "bias"] = b


def _lstm_entries(sd, seed, prefix, dim, layers=2):
    bound = 1.0 / math.sqrt(dim)
    for l in range(layers):
        for nm, shape in (("weight_ih", (4 * dim, dim)), ("weight_hh", (4 * dim, dim)),
                          ("bias_ih", (4 * dim,)), ("bias_hh", (4 * dim,))):
            sd[f"{prefix}lstm.{nm}_l{l}"] = _uniform(seed, f"{prefix}{nm}{l}", shape, bound)


def _encoder_entries(sd, seed, prefix, cfg: CodecConfig):
    """SEANetEncoder layout (seanet.py:63-153) with n_residual_layers=1."""
    nf = cfg.n_filters
    _conv_entries(sd, seed, f"{prefix}model.0.conv.conv.", nf, cfg.channels, cfg.kernel_size)
    mult, idx = 1, 1
    for r in reversed(cfg.ratios):
        c = mult * nf
        _conv_entries(sd, seed, f"{prefix}model.{idx}.block.1.conv.conv.", c // cfg.compress, c, cfg.residual_kernel_size)
        _conv_entries(sd, seed, f"{prefix}model.{idx}.block.3.conv.conv.", c, c // cfg.compress, 1)
        _conv_entries(sd, seed, f"{prefix}model.{idx + 2}.conv.conv.", 2 * c, c, 2 * r)
        idx += 3
        mult *= 2
    _lstm_entries(sd, seed, f"{prefix}model.{idx}.", mult * nf, cfg.lstm)
    _conv_entries(sd, seed, f"{prefix}model.{idx + 2}.conv.conv.", cfg.dimension, mult * nf, cfg.last_kernel_size)


def _decoder_entries(sd, seed, prefix, cfg: CodecConfig):
    """SEANetDecoder layout (seanet.py:156-258) with n_residual_layers=1."""
    nf = cfg.n_filters
    mult = 2 ** len(cfg.ratios)
    _conv_entries(sd, seed, f"{prefix}model.0.conv.conv.", mult * nf, cfg.dimension, cfg.kernel_size)
    _lstm_entries(sd, seed, f"{prefix}model.1.", mult * nf, cfg.lstm)
    idx = 3
    for r in cfg.ratios:
        c = mult * nf
        _conv_entries(sd, seed, f"{prefix}model.{idx}.convtr.convtr.", c // 2, c, 2 * r, transpose=True)
        h = c // 2
        _conv_entries(sd, seed, f"{prefix}model.{idx + 1}.block.1.conv.conv.", h // cfg.compress, h, cfg.residual_kernel_size)
        _conv_entries(sd, seed, f"{prefix}model.{idx + 1}.block.3.conv.conv.", h, h // cfg.compress, 1)
        idx += 3
        mult //= 2
    _conv_entries(sd, seed, f"{prefix}model.{idx}.conv.conv.", cfg.channels, nf, cfg.last_kernel_size)


def make_codebook(cfg: CodecConfig, seed: int, q: int, mu: torch.Tensor, sigma: torch.Tensor, rho: float = 0.5):
    """Stage-q RVQ codebook = mu + sigma * (seeded random directions of radius rho*sqrt(dim)).

    `kmeans_init` leaves the reference's codebooks at zeros (core_vq.py:115-116) and isotropic N(0,s)
    codebooks are degenerate for random-init latents (the minimum-norm code always wins).
    Equal-radius directions around the per-dimension mean/std of the stage's residual (measured on a
    calibration batch, `calibrate_codebooks`) give a few hundred distinct indices per codebook."""
    z = _normal(seed, f"codebook{q}", (cfg.bins, cfg.dimension))
    z = z / z.norm(dim=1, keepdim=True) * (math.sqrt(cfg.dimension) * rho)
    return (mu.view(1, -1).float() + sigma.view(1, -1).float() * z).contiguous()


def make_codec_state_dict(cfg: CodecConfig = CodecConfig(), seed: int = 0, codebook_mu=None,
                          codebook_sigma=None) -> "OrderedDict[str, torch.Tensor]":
    """state_dict for reference `WMEncodecModel` (keys: SURVEY Appendix D).
    codebook_mu / codebook_sigma: [n_q, dimension] calibration statistics (see make_codebook)."""
    sd = OrderedDict()
    _encoder_entries(sd, seed, "encoder.", cfg)
    _decoder_entries(sd, seed, "decoder.", cfg)
    if codebook_mu is None:
        codebook_mu = torch.zeros(cfg.n_q, cfg.dimension)
    if codebook_sigma is None:
        codebook_sigma = torch.full((cfg.n_q, cfg.dimension), 0.01)
    codebook_mu = torch.as_tensor(codebook_mu, dtype=torch.float32)
    codebook_sigma = torch.as_tensor(codebook_sigma, dtype=torch.float32)
    for q in range(cfg.n_q):
        p = f"quantizer.vq.layers.{q}._codebook."
        emb = make_codebook(cfg, seed, q, codebook_mu[q], codebook_sigma[q])
        sd[p + "inited"] = torch.tensor([1.0])
        sd[p + "cluster_size"] = torch.ones(cfg.bins)
        sd[p + "embed"] = emb
        sd[p + "embed_avg"] = emb.clone()
    _decoder_entries(sd, seed + 101, "wmdecoder.", cfg)
    _encoder_entries(sd, seed + 202, "wmdecoder.skip_encoder.", cfg)
    _encoder_entries(sd, seed + 303, "wmdecoder.wm_encoder.", cfg)
    e = cfg.dimension // 16
    sd["wmdecoder.wm_embed.weight"] = _normal(seed, "wm_embed", (2, e), 0.6)   # some rows exceed max_norm=1
    nf = cfg.n_filters
    mult = 2 ** len(cfg.ratios)
    chans = [cfg.dimension, mult // 2 * nf, mult // 4 * nf, mult // 8 * nf]
    for i, c in enumerate(chans):
        _conv_entries(sd, seed, f"wmdecoder.wm_proj{i}.1.conv.conv.", c, c + e, 1, weight_norm=False)
    _conv_entries(sd, seed, "wmdecoder.wm_predictor.1.conv.conv.", 2, cfg.dimension, 1, weight_norm=False)
    return sd


def calibrate_codebooks(cfg: CodecConfig, seed: int, latents: torch.Tensor):
    """latents [N, dimension] (encoder outputs of a calibration batch, any implementation).
    Returns (mu [n_q, dim], sigma [n_q, dim]) for `make_codec_state_dict` by running the RVQ
    residual recursion (core_vq.py:382-392) with the recipe codebooks."""
    res = latents.detach().float().cpu().clone()
    mus, sgs = [], []
    for q in range(cfg.n_q):
        mu, sg = res.mean(0), res.std(0)
        mus.append(mu)
        sgs.append(sg)
        E = make_codebook(cfg, seed, q, mu, sg)
        d = -(res.pow(2).sum(1, keepdim=True) - 2 * res @ E.t() + E.t().pow(2).sum(0, keepdim=True))
        res = res - E[d.max(-1).indices]
    return torch.stack(mus), torch.stack(sgs)
