"""TEST INFRASTRUCTURE — pins the oracle against the UNMODIFIED reference and writes golden fixtures.

Run in the build container (needs /root/reference; does not exist on the GPU box):

    python oracle/gen_golden.py            # writes tests/golden/*.npz, prints the oracle-vs-reference report

For every case the reference's own code (`models/ssr.py::SSR_Speech.inference`, WM-Encodec
`encode/decode/wmdecode`) is executed on CPU with a seeded synthetic checkpoint
(`ssr_speech_b200.synth`), the oracle restatement is executed on the same inputs, the two are
compared, and the reference's outputs are stored.  Fixtures hold inputs + reference outputs only
(weights are regenerated from their seed by the tests).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_loader  # noqa: E402
from lm_oracle import LMOracle  # noqa: E402
from codec_oracle import CodecOracle  # noqa: E402
import ssr_speech_b200 as pkg  # noqa: E402
from ssr_speech_b200 import seq as seqmod  # noqa: E402
from ssr_speech_b200.config import CodecConfig, cfg_tiny  # noqa: E402
from ssr_speech_b200.synth import calibrate_codebooks, make_codec_state_dict, make_lm_state_dict  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

LM_CASES = [
    # name, Lx, T, mask_interval, decode kwargs, seed
    dict(name="tts_greedy", Lx=7, T=20, spans=[[20, 20]], seed=11,
         kw=dict(top_k=1, top_p=1.0, temperature=1.0, stop_repetition=-1, kvcache=1, cfg_coef=1.5, cfg_stride=1, aug_text=False)),
    dict(name="edit_cfg_sampled", Lx=9, T=40, spans=[[12, 25]], seed=12,
         kw=dict(top_k=0, top_p=0.8, temperature=1.0, stop_repetition=2, kvcache=1, cfg_coef=1.5, cfg_stride=2, aug_text=True)),
    dict(name="edit2_cfg_greedy", Lx=8, T=36, spans=[[5, 9], [20, 28]], seed=13,
         kw=dict(top_k=1, top_p=1.0, temperature=1.0, stop_repetition=-1, kvcache=1, cfg_coef=2.0, cfg_stride=3, aug_text=True)),
    dict(name="tts_cfg_temp_topk", Lx=6, T=16, spans=[[16, 16]], seed=14,
         kw=dict(top_k=20, top_p=0.9, temperature=0.7, stop_repetition=3, kvcache=1, cfg_coef=1.5, cfg_stride=1, aug_text=True)),
    dict(name="edit_head_nokv", Lx=6, T=24, spans=[[0, 6]], seed=15,
         kw=dict(top_k=1, top_p=1.0, temperature=1.0, stop_repetition=-1, kvcache=0, cfg_coef=1.5, cfg_stride=1, aug_text=False)),
    # three spans (max_n_spans), the middle one an insertion (empty interval)
    dict(name="edit3_cfg_sampled", Lx=10, T=44, spans=[[3, 8], [15, 15], [30, 36]], seed=16,
         kw=dict(top_k=0, top_p=0.9, temperature=1.0, stop_repetition=2, kvcache=1, cfg_coef=1.5, cfg_stride=2, aug_text=True)),
    # aug_context (ssr.py:564-593): prompt text/codes are prepended, intervals shift by out_len, outputs shift back
    dict(name="ctx_edit_greedy", Lx=7, T=30, spans=[[10, 18]], seed=17, ctx=dict(Lpx=5, Tp=12),
         kw=dict(top_k=1, top_p=1.0, temperature=1.0, stop_repetition=-1, kvcache=1, cfg_coef=1.5, cfg_stride=1, aug_text=False,
                 aug_context=True)),
    dict(name="ctx_tts_cfg_sampled", Lx=6, T=18, spans=[[18, 18]], seed=18, ctx=dict(Lpx=4, Tp=9),
         kw=dict(top_k=0, top_p=0.8, temperature=1.0, stop_repetition=2, kvcache=1, cfg_coef=1.5, cfg_stride=2, aug_text=True,
                 aug_context=True)),
    # stop_repetition penalty (ssr.py:727-736): the greedy roll-outs of these seeds sit on tokens 34 / 25 for tens of steps, so with
    # those ids as "silence" tokens the consecutive-repeat counter exceeds stop_repetition and the logit is rescaled
    dict(name="tts_rep_greedy", Lx=12, T=20, spans=[[20, 20]], seed=26, silence=[25, 34, 37],
         kw=dict(top_k=1, top_p=1.0, temperature=1.0, stop_repetition=1, kvcache=1, cfg_coef=1.5, cfg_stride=1, aug_text=False)),
    dict(name="tts_rep_cfg_lowtemp", Lx=12, T=20, spans=[[20, 20]], seed=26, silence=[25, 34, 37, 42],
         kw=dict(top_k=3, top_p=1.0, temperature=0.3, stop_repetition=1, kvcache=1, cfg_coef=1.5, cfg_stride=2, aug_text=True)),
    dict(name="tts_rep_cfg_greedy", Lx=12, T=20, spans=[[20, 20]], seed=27, silence=[25, 34, 37, 42],
         kw=dict(top_k=1, top_p=1.0, temperature=1.0, stop_repetition=2, kvcache=1, cfg_coef=1.5, cfg_stride=1, aug_text=True)),
]
SILENCE = [3, 17, 40]   # tiny-vocab stand-ins for the reference default [1388,1898,131]


def run_lm_cases(only=None):
    ssr = ref_loader.load_reference_ssr()
    cfg = cfg_tiny()
    sd = make_lm_state_dict(cfg, seed=7)
    model = ssr.SSR_Speech(cfg.to_namespace()).eval()
    missing = model.load_state_dict(sd, strict=True)
    print("reference SSR_Speech loaded synthetic state_dict:", missing)
    oracle = LMOracle(cfg, sd)
    report = []
    for case in LM_CASES:
        if only and case["name"] not in only:
            continue
        g = torch.Generator().manual_seed(case["seed"])
        Lx, T = case["Lx"], case["T"]
        x = torch.randint(0, cfg.text_vocab_size, (1, Lx), generator=g)
        y = torch.randint(0, cfg.audio_vocab_size, (1, T, cfg.n_codebooks), generator=g)
        mi = torch.tensor([case["spans"]], dtype=torch.long)
        kw = dict(case["kw"])
        SILENCE = case.get("silence", globals()["SILENCE"])
        ctx = case.get("ctx")
        if ctx:
            px = torch.randint(0, cfg.text_vocab_size, (1, ctx["Lpx"]), generator=g)
            pr = torch.randint(0, cfg.audio_vocab_size, (1, ctx["Tp"], cfg.n_codebooks), generator=g)
        else:
            px, pr = x, y
        # ---- reference run (global CPU RNG seeded like inference_v2.seed_everything) -----------
        torch.manual_seed(case["seed"])
        with torch.no_grad():
            res, marks, masks, nmi = model.inference(x, torch.tensor([Lx]), px, torch.tensor([px.shape[1]]), y, pr,
                                                     mask_interval=mi, silence_tokens=SILENCE, **kw)
        # ---- oracle run with the same RNG stream ------------------------------------------------
        use_ctx = bool(kw.get("aug_context")) and sum(b - a for a, b in case["spans"]) < 2 * 50      # ssr.py:564-568
        xo = torch.cat([px[0], x[0]]) if use_ctx else x[0]                                           # ssr.py:583,592
        yo = np.concatenate([pr[0].T.numpy(), y[0].T.numpy()], 1) if use_ctx else y[0].T.numpy().copy()
        prep = seqmod.prepare(cfg, yo, case["spans"], out_len=pr.shape[1] if use_ctx else 0)
        okw = {k: v for k, v in kw.items() if k not in ("kvcache", "aug_context")}
        torch.manual_seed(case["seed"])
        spans = oracle.inference(xo, torch.from_numpy(prep.prompt_tokens), prep.num_spans,
                                 silence_tokens=SILENCE, incremental=bool(kw["kvcache"]), **okw)
        ores, omarks, omasks, onmi = seqmod.finalize(cfg, prep, spans)
        same = (ores.shape == tuple(res[0].shape) and np.array_equal(ores, res[0].numpy())
                and np.array_equal(omarks, marks[0].numpy()) and omasks == [tuple(m) for m in masks]
                and onmi == [tuple(m) for m in nmi])
        # ---- the same run again through explicit exponential noise (what the GPU path consumes) --
        torch.manual_seed(case["seed"])
        uncond = torch.randint(0, cfg.n_text_tokens, (1, xo.shape[0]))[0] if kw["aug_text"] else None
        n_steps = sum(len(s) for s in spans)
        noise = torch.stack([torch.empty(cfg.n_codebooks, cfg.n_audio_tokens).exponential_(1) for _ in range(n_steps)])
        trace = []
        spans2 = oracle.inference(xo, torch.from_numpy(prep.prompt_tokens), prep.num_spans, silence_tokens=SILENCE,
                                  uncond_x=uncond, noise=noise, trace=trace, **okw)
        same_noise = all(np.array_equal(a, b) for a, b in zip(spans, spans2))
        report.append((case["name"], same, same_noise, tuple(res.shape)))
        print(f"[lm] {case['name']:<20} oracle==reference: {same}   noise-path==multinomial-path: {same_noise}   res {tuple(res.shape)}")
        np.savez_compressed(
            os.path.join(GOLD, f"lm_{case['name']}.npz"),
            x=x[0].numpy(), y=y[0].numpy(), mask_interval=np.asarray(case["spans"]), silence=np.asarray(SILENCE),
            kw=json.dumps(kw), seed=case["seed"], weights_seed=7,
            uncond_x=(uncond.numpy() if uncond is not None else np.zeros(0, np.int64)),
            prompt_x=(px[0].numpy() if ctx else np.zeros(0, np.int64)),
            prompt=(pr[0].numpy() if ctx else np.zeros((0, cfg.n_codebooks), np.int64)),
            noise=noise.numpy().astype(np.float32),
            ref_res=res[0].numpy(), ref_marks=marks[0].numpy(), ref_masks=np.asarray(masks), ref_nmi=np.asarray(nmi),
            ref_span_lens=np.asarray([len(s) for s in spans]),
            ref_span_tokens=np.concatenate(spans, 0),
            raw_logits_step0=trace[0].raw_logits.numpy(), probs_step0=trace[0].probs.numpy(),
            raw_logits_last=trace[-1].raw_logits.numpy(),
        )
    if only:
        return report, None
    # teacher-forced logits fixture straight from the reference modules (no loop)
    g = torch.Generator().manual_seed(99)
    x = torch.randint(0, cfg.text_vocab_size, (11,), generator=g)
    toks = torch.randint(0, cfg.audio_vocab_size, (cfg.n_codebooks, 30), generator=g)
    with torch.no_grad():
        xi = model.text_positional_embedding(model.text_embedding(x[None]))
        yi = model.audio_positional_embedding(model.embed_y(toks[:, :, None]))
        S = xi.shape[1] + yi.shape[1]
        mask = torch.triu(torch.ones(S, S), diagonal=1).bool()
        out, _ = model.decoder((torch.cat([xi, yi], 1), None), mask=mask)
        h = out[:, xi.shape[1]:]
        ref_logits = torch.stack([model.predict_layer[k](h) for k in range(cfg.n_codebooks)], 2)[0]  # [T,K,V]
    ol = oracle.teacher_forced_logits(x, toks)
    err = (ol - ref_logits).abs().max().item()
    print(f"[lm] teacher-forced logits oracle vs reference max-abs err {err:.3e}")
    np.savez_compressed(os.path.join(GOLD, "lm_teacher_forced.npz"), x=x.numpy(), toks=toks.numpy(),
                        ref_logits=ref_logits.numpy(), weights_seed=7)
    return report, err


def run_codec_cases():
    cfg = CodecConfig()
    model = ref_loader.build_reference_codec()
    # calibration pass: per-stage residual statistics for the codebook recipe (synth.make_codebook)
    sd = make_codec_state_dict(cfg, seed=3)
    model.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(1234)
    t = torch.arange(6400) / 16000.0
    wav = torch.stack([0.1 * torch.randn(6400, generator=g) * (0.3 + torch.sin(2 * np.pi * (3 + 2 * i) * t) ** 2)
                       + 0.05 * torch.sin(2 * np.pi * (200 + 150 * i) * t) for i in range(2)])[:, None]
    wav[1, :, 4000:] *= 0.2
    tc = torch.arange(32000) / 16000.0
    cal = torch.stack([0.1 * torch.randn(32000, generator=g) * (0.3 + torch.sin(2 * np.pi * (1 + i) * tc) ** 2)
                       + 0.05 * torch.sin(2 * np.pi * (200 + 150 * i) * tc) for i in range(4)])[:, None]
    with torch.no_grad():
        emb = model.encoder(cal)
    mu, sigma = calibrate_codebooks(cfg, 3, emb.permute(0, 2, 1).reshape(-1, cfg.dimension))
    sd = make_codec_state_dict(cfg, seed=3, codebook_mu=mu, codebook_sigma=sigma)
    model.load_state_dict(sd, strict=True)
    oracle = CodecOracle(cfg, sd)
    with torch.no_grad():
        codes, scale, emb = model.encode(wav)
        dec = model.decode(codes, None)
        marks = torch.zeros(2, codes.shape[-1], dtype=torch.long)
        marks[0, 5:12] = 1
        marks[1, 0:4] = 1
        wm, mlog = model.wmdecode(codes, marks, wav, None)
    ocodes, _, oemb = oracle.encode(wav)
    odec = oracle.decode(codes)
    owm, omlog = oracle.wmdecode(codes, marks, wav)
    rep = dict(
        emb_err=(oemb - emb).abs().max().item(), emb_max=emb.abs().max().item(),
        codes_equal=bool(torch.equal(ocodes, codes)), n_distinct=[int(codes[:, q].unique().numel()) for q in range(cfg.n_q)],
        codes_given_ref_emb=bool(torch.equal(oracle.rvq_encode(emb), codes)),
        dec_err=(odec - dec).abs().max().item(), dec_max=dec.abs().max().item(),
        wm_err=(owm - wm).abs().max().item(), wm_max=wm.abs().max().item(),
        mlog_err=(omlog - mlog).abs().max().item(),
    )
    print("[codec] oracle vs reference:", json.dumps(rep))
    np.savez_compressed(os.path.join(GOLD, "codec_small.npz"), wav=wav.numpy(), weights_seed=3,
                        codebook_mu=mu.numpy(), codebook_sigma=sigma.numpy(), ref_emb=emb.numpy(), ref_codes=codes.numpy(),
                        ref_dec=dec.numpy(), marks=marks.numpy(), ref_wm=wm.numpy(), ref_mark_logits=mlog.numpy())
    # config 1 of BASELINE.json: demo wav round trip, full length (126 880 samples -> 127 040 -> 397 frames)
    try:
        from scipy.io import wavfile
        sr, data = wavfile.read(os.path.join(ref_loader.REF_ROOT, "demo", "84_121550_000074_000000.wav"))
        data = torch.from_numpy(np.asarray(data, dtype=np.float32))[None, None]
        pad = (320 - data.shape[-1] % 320) % 320          # data/tokenizer.py:148-151
        data = torch.nn.functional.pad(data, (0, pad))
        with torch.no_grad():
            c_full, _, e_full = model.encode(data)
            d_full = model.decode(c_full, None)
        oc, _, oe = oracle.encode(data)
        od = oracle.decode(c_full)
        print(f"[codec] demo wav sr={sr} len={data.shape[-1]} frames={c_full.shape[-1]} "
              f"codes_equal={bool(torch.equal(oc, c_full))} emb_err={(oe - e_full).abs().max().item():.3e} "
              f"dec_err={(od - d_full).abs().max().item():.3e}")
        # BASELINE configs[0]: the WHOLE demo utterance (397 frames), decode and wmdecode with marks[100:200] = 1
        marks_full = torch.zeros(1, c_full.shape[-1], dtype=torch.long)
        marks_full[0, 100:200] = 1
        with torch.no_grad():
            wm_full, mlog_full = model.wmdecode(c_full, marks_full, data, None)
        owm, omlog = oracle.wmdecode(c_full, marks_full, data)
        print(f"[codec] demo wav wmdecode oracle vs reference: wav err {(owm - wm_full).abs().max().item():.3e} "
              f"mark-logit err {(omlog - mlog_full).abs().max().item():.3e}")
        np.savez_compressed(os.path.join(GOLD, "codec_demo.npz"), wav=data.numpy(), weights_seed=3,
                            codebook_mu=mu.numpy(), codebook_sigma=sigma.numpy(), ref_emb=e_full.numpy(), ref_codes=c_full.numpy(),
                            ref_dec=d_full.numpy(), marks=marks_full.numpy(), ref_wm=wm_full.numpy(),
                            ref_mark_logits=mlog_full.numpy(), n_samples=int(data.shape[-1] - pad))
    except Exception as e:  # pragma: no cover
        print("[codec] demo wav case skipped:", repr(e))
    return rep


def synth_train_batch(cfg, seed: int, lens=((9, 26), (12, 33), (7, 19))):
    """A dataset-style training batch for SSR_Speech.forward (models/ssr.py:280-379): per utterance [kept prefix, <mts>, kept
    suffix, <mts>, masked span, <eog>] in the delayed codebook pattern (codebook k shifted right by k, empty_token fill), padded
    with audio_pad_token / text_pad_token to the batch maxima.  Exercises every token class the loss masks look at."""
    g = torch.Generator().manual_seed(seed)
    K, B = cfg.n_codebooks, len(lens)
    xs, ys = [], []
    for Lx, T in lens:
        xs.append(torch.randint(0, cfg.text_vocab_size, (Lx,), generator=g))
        a = int(torch.randint(2, T // 2, (1,), generator=g))
        b = int(torch.randint(a + 2, T - 2, (1,), generator=g))
        tok = torch.randint(0, cfg.audio_vocab_size, (K, T), generator=g)
        mts = torch.full((K, 1), cfg.mts, dtype=torch.long)
        segs = [tok[:, :a], mts, tok[:, b:], mts, torch.cat([tok[:, a:b], torch.full((K, 1), cfg.eog, dtype=torch.long)], 1)]
        flat = torch.cat(segs, 1)
        L = flat.shape[1] + K - 1
        y = torch.full((K, L), cfg.empty_token, dtype=torch.long)
        for k in range(K):
            y[k, k:k + flat.shape[1]] = flat[k]
        ys.append(y)
    x_lens = torch.tensor([x.shape[0] for x in xs])
    y_lens = torch.tensor([y.shape[1] for y in ys])
    x = torch.full((B, int(x_lens.max())), cfg.text_pad_token, dtype=torch.long)
    y = torch.full((B, K, int(y_lens.max())), cfg.audio_pad_token, dtype=torch.long)
    for i in range(B):
        x[i, :x_lens[i]] = xs[i]
        y[i, :, :y_lens[i]] = ys[i]
    return x, x_lens, y, y_lens


def run_train_forward_cases():
    """the unmodified reference's training forward (eval mode: dropout inactive) on a synthetic batch, for both
    settings of the loss-mask switches; the oracle's forward_loss on the same batch; fixtures for the CPU and GPU tests."""
    ssr = ref_loader.load_reference_ssr()
    cfg = cfg_tiny()
    sd = make_lm_state_dict(cfg, seed=7)
    ns = cfg.to_namespace()
    model = ssr.SSR_Speech(ns).eval()
    model.load_state_dict(sd, strict=True)
    oracle = LMOracle(cfg, sd)
    out = {}
    for tag, (pmt, pall, cw) in {"default": (1, 0, "[5,1,0.5,0.1]"), "all": (0, 1, None)}.items():
        x, x_lens, y, y_lens = synth_train_batch(cfg, seed=41)
        model.args.predict_mask_token, model.args.predict_all, model.args.codebook_weight = pmt, pall, cw
        with torch.no_grad():
            ref = model({"x": x, "x_lens": x_lens, "y": y, "y_lens": y_lens})
        got = oracle.forward_loss(x, x_lens, y, y_lens, predict_mask_token=bool(pmt), predict_all=bool(pall),
                                  codebook_weight=None if cw is None else eval(cw))
        rl, rt, rn = float(ref["loss"]), float(ref["top10acc"]), int(ref["effective_ntoken"])
        rb = [float(v) for v in ref["top10acc_by_codebook"]]
        print(f"[train] {tag:<8} reference loss {rl:.6f} top10acc {rt:.4f} ntoken {rn} | oracle loss {got['loss']:.6f} "
              f"top10acc {got['top10acc']:.4f} ntoken {got['effective_ntoken']} | rel.err {abs(got['loss'] - rl) / abs(rl):.2e}")
        out.update({f"{tag}_loss": rl, f"{tag}_top10acc": rt, f"{tag}_ntoken": rn, f"{tag}_top10acc_by_codebook": np.asarray(rb),
                    f"{tag}_predict_mask_token": pmt, f"{tag}_predict_all": pall,
                    f"{tag}_codebook_weight": np.asarray(eval(cw) if cw else [1.0] * cfg.n_codebooks, dtype=np.float64)})
    np.savez_compressed(os.path.join(GOLD, "lm_train_forward.npz"), x=x.numpy(), x_lens=x_lens.numpy(), y=y.numpy(),
                        y_lens=y_lens.numpy(), weights_seed=7, **out)


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    which = sys.argv[1] if len(sys.argv) > 1 else "all"       # all | lm | codec | train | lm:<case>,<case> (only those fixtures)
    if which.startswith("lm:"):
        run_lm_cases(only=set(which[3:].split(",")))
    if which in ("all", "lm"):
        run_lm_cases()
    if which in ("all", "codec"):
        run_codec_cases()
    if which in ("all", "train"):
        run_train_forward_cases()
