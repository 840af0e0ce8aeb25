"""`SSR_Speech` — drop-in for the inference path of reference models/ssr.py::SSR_Speech.

Same constructor (`SSR_Speech(args)` with the Namespace stored in ``ckpt["config"]``), same
``load_state_dict`` keys, same ``inference(...)`` signature, assertions and
return values (models/ssr.py:504-524, :552-561, :804-812).  Host-side sequence surgery lives in
`seq.py`; everything between the prologue and the epilogue — embeddings, 16 decoder layers with an
in-place KV cache, prediction heads, CFG, logit rules, top-k/top-p sampling and the EOG/span state
machine — runs inside libssr_b200.so as sm_100a kernels replayed from a CUDA graph.

Extension over the reference (which asserts batch 1, ssr.py:559-561): `inference_batch` decodes many
utterances at once; utterance i behaves exactly like an independent `inference` call.
"""
from __future__ import annotations

import copy
import ctypes as C
import os
from argparse import Namespace
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib, seq
from .config import SSRConfig

_PE_MIN = 4000   # reference starts with a 4000-position table and auto-extends (embedding.py:67-92)


def sinusoid_table(n_pos: int, dim: int) -> torch.Tensor:
    """Same construction as models/modules/embedding.py:76-92 (fp32 torch CPU ops)."""
    import math
    pos = torch.arange(0, n_pos, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, dim, 2, dtype=torch.float32) * -(math.log(10000.0) / dim))
    pe = torch.zeros(n_pos, dim)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


class SSR_Speech:
    def __init__(self, args: Optional[Namespace] = None, config: Optional[Dict] = None, precision: Optional[str] = None,
                 gemm_impl: int = 0):
        if args is None:
            if config is None:
                raise ValueError("Either `args` or `config` must be provided.")
            args = Namespace(**config)
        elif config is not None:
            raise ValueError("Cannot provide both `args` and `config`.")
        self.args = copy.copy(args)
        self.cfg = SSRConfig.from_args(args)
        # mirror the attribute normalisation of ssr.py:113-119
        self.args.n_special = self.cfg.n_special
        self.args.eos = self.cfg.eos
        self.args.audio_vocab_size = self.cfg.audio_vocab_size
        self.n_text_tokens = self.cfg.n_text_tokens
        self.n_audio_tokens = [self.cfg.n_audio_tokens] * self.cfg.n_codebooks
        precision = precision or os.environ.get("SSRB_PRECISION", "bf16")
        assert precision in ("bf16", "fp32"), precision
        self.precision = precision
        self.gemm_impl = int(os.environ.get("SSRB_GEMM_IMPL", gemm_impl))
        self._sd: Optional[Dict[str, torch.Tensor]] = None
        self._device: Optional[torch.device] = None
        self._h = None           # ssrb_lm*
        self._cap = None         # (max_rows, max prompt positions s0, max_prefill_tokens, max_steps)
        self.training = False
        self.last_stats: Dict[str, float] = {}
        self.poll_every = 16     # decode iterations enqueued between two `done` polls

    # ---- nn.Module-like surface used by inference_v2.py:198-204 -------------------------------------
    def load_state_dict(self, state_dict, strict: bool = True):
        sd = {k: v.detach().to("cpu", torch.float32).contiguous() for k, v in state_dict.items() if torch.is_tensor(v)}
        need = ["text_embedding.word_embeddings.weight", "decoder.norm.weight", "predict_layer.0.0.weight"]
        missing = [k for k in need if k not in sd]
        if missing and strict:
            raise RuntimeError(f"Missing key(s) in state_dict: {missing}")
        self._sd = sd
        self._destroy()
        return torch.nn.modules.module._IncompatibleKeys(missing, [])

    def state_dict(self):
        return dict(self._sd or {})

    def to(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("ssr_speech_b200.SSR_Speech runs on CUDA devices only (no CPU fallback)")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        if self._device != device:
            self._destroy()
        self._device = device
        return self

    def cuda(self, index=None):
        return self.to(torch.device("cuda", index if index is not None else torch.cuda.current_device()))

    def eval(self):
        self.training = False
        return self

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    def _destroy(self):
        if self._h is not None:
            _lib.load().ssrb_lm_destroy(self._h)
            self._h = None
            self._cap = None

    # ---- engine management ----------------------------------------------------------------------------
    def _ensure_engine(self, rows: int, s0: int, prefill_tokens: int, max_steps: int):
        """Engine capacity: `rows` transformer rows, prompts of up to `s0` positions (text + audio + <mts>), `max_steps` loop
        iterations per utterance; the KV rows hold s0 + max_steps (+ slack) positions.  Grow-only: a rebuild keeps the
        largest value of every dimension seen so far (max_seq is derived, so growing max_steps also grows the cache)."""
        if self._sd is None:
            raise RuntimeError("load_state_dict must be called before inference")
        if self._device is None:
            raise RuntimeError("call .to('cuda') before inference (no CPU fallback)")
        cap = self._cap
        if self._h is not None and cap[0] >= rows and cap[1] >= s0 and cap[2] >= prefill_tokens and cap[3] >= max_steps:
            return
        if cap is not None:
            # grow-only, with headroom on the dimensions that outgrew the engine: a rebuild re-uploads every weight (3.3 GB for
            # the 830M model), so a per-utterance loop over slowly growing requests (inference_v2.py:331-333) must not rebuild
            # on every longer request.  The first build is exact; `reserve()` pre-sizes.
            def grow(need, have):
                return have if need <= have else max(need, (3 * have + 1) // 2)
            rows, s0 = max(rows, cap[0]), grow(s0, cap[1])
            prefill_tokens, max_steps = grow(prefill_tokens, cap[2]), grow(max_steps, cap[3])
        max_seq = s0 + max_steps + 8
        self._destroy()
        lib = _lib.load()
        c = self.cfg
        conf = _lib.LMConfig(
            d_model=c.d_model, n_head=c.nhead, n_layer=c.num_decoder_layers, ffn_dim=c.ffn_dim, n_codebooks=c.n_codebooks,
            n_audio_tokens=c.n_audio_tokens, n_text_tokens=c.n_text_tokens, head_hidden=c.head_hidden,
            empty_token=c.empty_token, eog=c.eog, eos=c.eos, sos=c.sos, mts=c.mts, max_n_spans=c.max_n_spans,
            max_rows=rows, max_seq=max_seq, max_prefill_tokens=prefill_tokens, max_steps=max_steps,
            weight_dtype=_lib.SSRB_DTYPE_BF16 if self.precision == "bf16" else _lib.SSRB_DTYPE_F32,
            gemm_impl=self.gemm_impl)
        h = C.c_void_p()
        with torch.cuda.device(self._device):
            _lib.check(lib.ssrb_lm_create(C.byref(conf), self._device.index, C.byref(h)), "ssrb_lm_create")
            self._h = h
            _lib.load_state_dict_into(lib.ssrb_lm_load_tensor, h, self._sd)
            n_pos = max(_PE_MIN, max_seq + 8)
            _lib.load_state_dict_into(lib.ssrb_lm_load_tensor, h, {"pe_table": sinusoid_table(n_pos, c.d_model)})
            _lib.check(lib.ssrb_lm_check_loaded(h), "ssrb_lm_check_loaded")
        self._cap = (rows, s0, prefill_tokens, max_steps)

    def reserve(self, n_utt: int, text_len: int, prompt_frames: int, aug_text: bool = True, n_spans: int = 1):
        """Optional: pre-build the engine for a given batch geometry (weights upload + allocations)."""
        rpu = 2 if aug_text else 1
        steps = self._max_steps(text_len, prompt_frames + 8, n_spans)
        s0 = text_len + prompt_frames + 8
        self._ensure_engine(n_utt * rpu, s0, min(max(s0, 16384), max(s0, s0 * n_utt * rpu)), steps)

    def _max_steps(self, x_len: int, prompt_len: int, n_spans: int) -> int:
        K = self.cfg.n_codebooks
        return n_spans * (max(10 * x_len - prompt_len, 0) + 2 + K) + 4

    # ---- the reference API ---------------------------------------------------------------------------------
    @torch.no_grad()
    def inference(self, x: torch.Tensor, x_lens: torch.Tensor, prompt_x: torch.Tensor, prompt_x_lens: torch.Tensor,
                  y: torch.Tensor, prompt: torch.Tensor, mask_interval: torch.Tensor, top_k: int = -100,
                  top_p: float = 1.0, temperature: float = 1.0, stop_repetition: int = -1, kvcache: int = 1,
                  silence_tokens: Sequence[int] = (1388, 1898, 131), cfg_coef: float = 1.5, cfg_stride: int = 1,
                  aug_text: bool = False, aug_context: bool = False, cfg_pretrained: bool = False,
                  _uncond_x: Optional[torch.Tensor] = None, _noise: Optional[torch.Tensor] = None):
        """Same contract as reference models/ssr.py:504-812 (batch 1).  `kvcache` is accepted for
        compatibility: the engine always decodes incrementally (the reference produces identical tokens
        either way).  `_uncond_x` / `_noise` are parity-test hooks."""
        assert cfg_coef >= 1.0, cfg_coef
        assert x.ndim == 2, x.shape
        assert x_lens.ndim == 1, x_lens.shape
        assert y.ndim == 3, y.shape
        assert prompt.ndim == 3, prompt.shape
        K = self.cfg.n_codebooks
        assert y.shape[0] == 1 and y.shape[2] == K, y.shape
        assert prompt.shape[0] == 1 and prompt.shape[2] == K, prompt.shape
        assert mask_interval.shape == torch.Size((1, mask_interval.shape[1], 2)), mask_interval
        out = self.inference_batch(
            [x[0]], [y[0]], [mask_interval[0]], prompt_xs=[prompt_x[0]], prompts=[prompt[0]], top_k=top_k, top_p=top_p,
            temperature=temperature, stop_repetition=stop_repetition, silence_tokens=silence_tokens, cfg_coef=cfg_coef,
            cfg_stride=cfg_stride, aug_text=aug_text, aug_context=aug_context, cfg_pretrained=cfg_pretrained,
            uncond_xs=None if _uncond_x is None else [_uncond_x], noise=_noise, device=y.device)
        return out[0]

    def _host_prepare(self, x, y, mask_interval, prompt_x, prompt, aug_text, aug_context, uncond_x):
        """Host-side prologue of one utterance (ssr.py:564-626): optional aug_context concatenation, span maths, delay
        pattern, <mts> insertion; draws the uncond text of the CFG row like the reference (global CPU generator)."""
        cfg, K = self.cfg, self.cfg.n_codebooks
        x = x.detach().to("cpu", torch.int64).reshape(-1)
        yk = y.detach().to("cpu", torch.int64)
        assert yk.ndim == 2 and yk.shape[1] == K, yk.shape
        y = yk.T.contiguous().numpy()                                    # [K, T]
        mi = torch.as_tensor(mask_interval).detach().to("cpu", torch.int64).reshape(-1, 2)
        context_len = int(sum(int(b) - int(a) for a, b in mi.tolist()))
        use_ctx = bool(aug_context and context_len < 2 * 50)             # ssr.py:564-568
        out_len = 0
        if use_ctx:
            p = prompt.detach().to("cpu", torch.int64).T.contiguous().numpy()
            px = prompt_x.detach().to("cpu", torch.int64).reshape(-1)
            out_len = p.shape[1]
            y = np.concatenate([p, y], axis=1)                           # ssr.py:581,591
            x = torch.cat([px, x], 0)                                    # ssr.py:583,592
        # nn.Embedding raises on ids outside its table (embedding.py:22-48); the device gathers would read stray memory
        if y.size and (y.min() < 0 or y.max() >= cfg.n_audio_tokens):
            raise IndexError("index out of range in self (audio token id outside the embedding table)")
        if x.numel() and (int(x.min()) < 0 or int(x.max()) >= cfg.n_text_tokens):
            raise IndexError("index out of range in self (phoneme id outside the embedding table)")
        prep = seq.prepare(cfg, y, mi.tolist(), out_len=out_len)
        ux = None
        if aug_text:
            if uncond_x is not None:
                ux = uncond_x.detach().to("cpu", torch.int64).reshape(-1)
            else:   # same draw as the reference: global CPU generator (ssr.py:574)
                ux = torch.randint(0, self.n_text_tokens, (1, x.shape[0]))[0]
            assert ux.shape[0] == x.shape[0]
            if int(ux.min()) < 0 or int(ux.max()) >= cfg.n_text_tokens:
                raise IndexError("index out of range in self (phoneme id outside the embedding table)")
        return prep, x, ux

    @torch.no_grad()
    def open_batch(self, xs: List[torch.Tensor], ys: List[torch.Tensor], mask_intervals: List, prompt_xs=None,
                   prompts=None, top_k: int = -100, top_p: float = 1.0, temperature: float = 1.0,
                   stop_repetition: int = -1, silence_tokens: Sequence[int] = (1388, 1898, 131),
                   cfg_coef: float = 1.5, cfg_stride: int = 1, aug_text: bool = False, aug_context: bool = False,
                   cfg_pretrained: bool = False, uncond_xs=None, noise: Optional[torch.Tensor] = None,
                   seed: Optional[int] = None, device=None):
        """Prologue of the loop: host sequence surgery, prefill of every row and the first sample (asynchronous)."""
        if cfg_pretrained:
            raise NotImplementedError("cfg_pretrained=True (key-padding on the uncond row, ssr.py:631-638) is not used "
                                      "by any reference entry point and is not implemented")
        assert cfg_coef >= 1.0, cfg_coef
        cfg = self.cfg
        K = cfg.n_codebooks
        U = len(xs)
        assert U == len(ys) == len(mask_intervals) and U > 0
        dev = torch.device(device) if device is not None else self._device
        if self._device is None:
            self.to(dev if dev is not None and dev.type == "cuda" else "cuda")
        rpu = 2 if aug_text else 1
        preps, rows_text, x_lens = [], [], []
        for i in range(U):
            prep, x, ux = self._host_prepare(xs[i], ys[i], mask_intervals[i], None if prompt_xs is None else prompt_xs[i],
                                             None if prompts is None else prompts[i], aug_text, aug_context,
                                             None if uncond_xs is None else uncond_xs[i])
            preps.append(prep)
            x_lens.append(int(x.shape[0]))
            rows_text.append(x)
            if aug_text:
                rows_text.append(ux)
        R = U * rpu
        Lmax = max(x_lens)
        Pmax = max(p.prompt_tokens.shape[1] for p in preps)
        steps = max(self._max_steps(x_lens[i], preps[i].prompt_tokens.shape[1], preps[i].num_spans) for i in range(U))
        s0s = [x_lens[i] + preps[i].prompt_tokens.shape[1] + 1 for i in range(U)]
        total_prefill = sum(s0s) * rpu
        self._ensure_engine(R, max(s0s), max(max(s0s), min(total_prefill, 16384)), steps)
        lib = _lib.load()
        text = np.zeros((R, Lmax), dtype=np.int32)
        for r, t in enumerate(rows_text):
            text[r, :t.shape[0]] = t.numpy()
        prom = np.zeros((U, K, max(Pmax, 1)), dtype=np.int32)
        for i, p in enumerate(preps):
            prom[i, :, :p.prompt_tokens.shape[1]] = p.prompt_tokens
        tl = np.asarray(x_lens, dtype=np.int32)
        pl = np.asarray([p.prompt_tokens.shape[1] for p in preps], dtype=np.int32)
        ns = np.asarray([p.num_spans for p in preps], dtype=np.int32)
        i32p = C.POINTER(C.c_int32)
        batch = _lib.LMBatch(n_utt=U, text=text.ctypes.data_as(i32p), text_stride=Lmax, text_len=tl.ctypes.data_as(i32p),
                             prompt=prom.ctypes.data_as(i32p), prompt_stride=prom.shape[2],
                             prompt_len=pl.ctypes.data_as(i32p), n_spans=ns.ctypes.data_as(i32p))
        sil = [int(t) for t in silence_tokens]
        if len(sil) > _lib.MAX_SILENCE:     # the reference accepts any list (ssr.py:727); truncating would change the sampling rule silently
            raise ValueError(f"at most {_lib.MAX_SILENCE} silence tokens are supported, got {len(sil)}")
        if seed is None:
            seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item()) if noise is None else 0
        sp = _lib.Sampling(top_k=int(top_k), top_p=float(top_p), temperature=float(temperature),
                           stop_repetition=int(stop_repetition), n_silence=len(sil),
                           silence_tokens=(C.c_int * _lib.MAX_SILENCE)(*(sil + [0] * (_lib.MAX_SILENCE - len(sil)))),
                           cfg_coef=float(cfg_coef), cfg_stride=int(cfg_stride), aug_text=int(bool(aug_text)), seed=seed)
        noise_dev = None
        if noise is not None:   # [n_steps, U, K, V] (or [n_steps, K, V] for one utterance); padded to max_steps
            nz = noise.detach().to(torch.float32)
            if nz.ndim == 3:
                nz = nz[:, None]
            assert nz.shape[1:] == (U, K, cfg.n_audio_tokens), nz.shape
            full = torch.ones(self._cap[3], U, K, cfg.n_audio_tokens, dtype=torch.float32)
            full[:min(nz.shape[0], self._cap[3])] = nz[:self._cap[3]]
            noise_dev = full.to(self._device).contiguous()
        self._keep = (text, prom, tl, pl, ns, noise_dev)      # host arrays must outlive the async begin
        with torch.cuda.device(self._device):
            st = _lib.stream_ptr()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            _lib.check(lib.ssrb_lm_begin(self._h, C.byref(batch), C.byref(sp),
                                         C.c_void_p(noise_dev.data_ptr()) if noise_dev is not None else None, st), "ssrb_lm_begin")
            ev1.record()
        # Exact upper bound on the loop iterations of this batch (the prefill's sample counts as iteration 1): the reference's
        # length guard (ssr.py:739) forces EOG at the latest after seq.expected_steps() iterations of a span, and a later span
        # starts from a longer sequence, so it can only be shorter.
        bound = max(preps[i].num_spans * seq.expected_steps(cfg, x_lens[i], preps[i].prompt_tokens.shape[1]) for i in range(U))
        return {"U": U, "preps": preps, "dev": dev, "ev0": ev0, "ev1": ev1, "bound": bound}

    @torch.no_grad()
    def inference_batch(self, xs, ys, mask_intervals, poll_every: Optional[int] = None, **kw):
        """xs[i]: [Lx_i] int64 phoneme ids; ys[i]: [T_i, K] int64 codes; mask_intervals[i]: [M_i, 2]; the remaining
        keyword arguments are those of `inference` (plus uncond_xs / noise / seed / device).
        Returns a list of (res [1,K,T_new] int64 (device), marks [1,T_new] int64 (CPU), masks, non_mask_intervals)."""
        ob = self.open_batch(xs, ys, mask_intervals, **kw)
        poll_every = int(poll_every or self.poll_every)
        lib = _lib.load()
        U, preps, dev = ob["U"], ob["preps"], ob["dev"]
        cfg, K = self.cfg, self.cfg.n_codebooks
        with torch.cuda.device(self._device):
            st = _lib.stream_ptr()
            ev2 = torch.cuda.Event(enable_timing=True)
            nd, it = C.c_int(0), C.c_int(0)
            while True:
                _lib.check(lib.ssrb_lm_poll(self._h, st, C.byref(nd), C.byref(it)), "ssrb_lm_poll")
                if nd.value >= U or it.value >= self._cap[3] + 1:
                    break
                # never enqueue past the last iteration the batch can need: an iteration after the last EOG still streams
                # every weight (8 of 513 iterations of the bench batch were such overshoot with a fixed chunk of 16)
                _lib.check(lib.ssrb_lm_decode(self._h, self._next_chunk(poll_every, ob["bound"], it.value), st), "ssrb_lm_decode")
            ev2.record()
            torch.cuda.synchronize()
            self.last_stats = {"prefill_ms": ob["ev0"].elapsed_time(ob["ev1"]), "decode_ms": ob["ev1"].elapsed_time(ev2),
                               "iterations": it.value}
            results = []
            for i in range(U):
                res, marks, masks, nmi = self._collect(i, preps[i])
                results.append((res.to(dev) if dev is not None else res, marks, masks, nmi))
        return results

    @staticmethod
    def _next_chunk(poll_every: int, bound: int, done_iterations: int) -> int:
        """Iterations to enqueue before the next poll: `poll_every`, but never past `bound` (the last iteration the batch can
        need); at least 1, so a bound that was too small could only slow the loop down, never end it early."""
        return int(min(poll_every, max(1, bound - done_iterations)))

    def _collect(self, slot: int, prep):
        """Reads the sampled tokens of one finished slot and runs the epilogue (ssr.py:774-812)."""
        lib, K = _lib.load(), self.cfg.n_codebooks
        buf = np.zeros((self._cap[3], K), dtype=np.int32)
        span_len = (C.c_int32 * _lib.MAX_SPANS)()
        n_tok = C.c_int(0)
        _lib.check(lib.ssrb_lm_read_tokens(self._h, _lib.stream_ptr(), slot, C.c_void_p(buf.ctypes.data), buf.shape[0],
                                           C.byref(n_tok), span_len), "ssrb_lm_read_tokens")
        toks = buf[:n_tok.value].astype(np.int64)
        spans, o = [], 0
        for s in range(prep.num_spans):
            spans.append(toks[o:o + span_len[s]])
            o += span_len[s]
        res, marks, masks, nmi = seq.finalize(self.cfg, prep, spans)
        return torch.from_numpy(res).unsqueeze(0).to(self._device), torch.from_numpy(marks).unsqueeze(0), masks, nmi

    @torch.no_grad()
    def serve(self, xs, ys, mask_intervals, max_slots: int, poll_every: Optional[int] = None, prompt_xs=None, prompts=None,
              uncond_xs=None, seed: Optional[int] = None, aug_text: bool = False, aug_context: bool = False, **kw):
        """Continuous batching: N requests through `max_slots` resident utterance slots.  The reference loops
        over utterances one at a time (inference_v2.py:331-333); here a slot whose utterance has emitted its last EOG is
        read back and refilled with the next request (ssrb_lm_admit: prefill into the slot's KV rows + first sample)
        while the other slots keep decoding.  Request i samples from Philox stream i whatever slot it lands in, so the
        result equals `inference_batch` over all N requests with the same seed (bit-exact in fp32 mode, where a row's
        arithmetic does not depend on its batch).  Returns the per-request tuples of `inference_batch`, in request order."""
        N = len(xs)
        assert N == len(ys) == len(mask_intervals) and N > 0 and max_slots > 0
        S = min(int(max_slots), N)
        rpu = 2 if aug_text else 1
        # host prologue of every request up front, in request order (the uncond-text draws consume the global CPU
        # generator in the same order as one big batch would)
        hp = [self._host_prepare(xs[i], ys[i], mask_intervals[i], None if prompt_xs is None else prompt_xs[i],
                                 None if prompts is None else prompts[i], aug_text, aug_context,
                                 None if uncond_xs is None else uncond_xs[i]) for i in range(N)]
        x_lens = [int(h[1].shape[0]) for h in hp]
        p_lens = [h[0].prompt_tokens.shape[1] for h in hp]
        steps = max(self._max_steps(x_lens[i], p_lens[i], hp[i][0].num_spans) for i in range(N))
        s0 = max(x_lens[i] + p_lens[i] + 1 for i in range(N))
        if self._device is None:
            self.to("cuda")
        self._ensure_engine(S * rpu, s0, max(s0, min(s0 * S * rpu, 16384)), steps)
        if seed is None:
            seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item())
        ob = self.open_batch(xs[:S], ys[:S], mask_intervals[:S], prompt_xs=None if prompt_xs is None else prompt_xs[:S],
                             prompts=None if prompts is None else prompts[:S], aug_text=aug_text, aug_context=aug_context,
                             uncond_xs=[h[2] for h in hp[:S]] if aug_text else None, seed=seed, **kw)
        lib = _lib.load()
        poll_every = int(poll_every or self.poll_every)
        results = [None] * N
        slot_req = list(range(S))           # request held by each slot (-1: empty)
        nxt, n_left = S, N
        flags = np.zeros(S, dtype=np.int32)
        it = C.c_int(0)
        limit = (N // S + 2) * (steps + 1) + 16
        with torch.cuda.device(self._device):
            st = _lib.stream_ptr()
            while n_left > 0:
                _lib.check(lib.ssrb_lm_poll_flags(self._h, st, C.c_void_p(flags.ctypes.data), C.byref(it)), "ssrb_lm_poll_flags")
                for slot in range(S):
                    r = slot_req[slot]
                    if r < 0 or not flags[slot]:
                        continue
                    results[r] = self._collect(slot, hp[r][0])
                    n_left -= 1
                    slot_req[slot] = -1
                    if nxt < N:
                        prep, x, ux = hp[nxt]
                        text = np.stack([x.numpy()] + ([ux.numpy()] if aug_text else [])).astype(np.int32)
                        prom = np.ascontiguousarray(prep.prompt_tokens, dtype=np.int32)
                        _lib.check(lib.ssrb_lm_admit(self._h, slot, C.c_void_p(text.ctypes.data), int(text.shape[1]),
                                                     C.c_void_p(prom.ctypes.data), int(prom.shape[1]), int(prep.num_spans),
                                                     int(nxt), st), "ssrb_lm_admit")
                        slot_req[slot] = nxt
                        nxt += 1
                if n_left == 0:
                    break
                if it.value > limit:
                    raise RuntimeError("serve: iteration limit exceeded")
                _lib.check(lib.ssrb_lm_decode(self._h, poll_every, st), "ssrb_lm_decode")
            torch.cuda.synchronize()
        self.last_stats = {"iterations": it.value}
        return results

    # ---- test hooks -------------------------------------------------------------------------------------------
    @torch.no_grad()
    def teacher_forced_logits(self, x: torch.Tensor, audio_tokens: torch.Tensor) -> torch.Tensor:
        """x [Lx], audio_tokens [K, Ty] -> fp32 logits [Ty, K, V] (CPU) via the prefill path."""
        K, V = self.cfg.n_codebooks, self.cfg.n_audio_tokens
        Lx, Ty = int(x.shape[0]), int(audio_tokens.shape[1])
        if self._device is None:
            self.to("cuda")
        self._ensure_engine(2, Lx + Ty + 8, Lx + Ty + 16, 8)
        xt = x.detach().to("cpu", torch.int32).contiguous().numpy()
        at = audio_tokens.detach().to("cpu", torch.int32).contiguous().numpy()
        out = np.zeros((Ty, K, V), dtype=np.float32)
        with torch.cuda.device(self._device):
            _lib.check(_lib.load().ssrb_lm_teacher_forced(self._h, C.c_void_p(xt.ctypes.data), Lx, C.c_void_p(at.ctypes.data),
                                                          Ty, C.c_void_p(out.ctypes.data), _lib.stream_ptr()), "teacher_forced")
        return torch.from_numpy(out)

    @torch.no_grad()
    def forward(self, batch):
        """Training forward / loss of the reference (models/ssr.py:280-379) in eval mode — forward only, no gradients:
        batch = {"x" [B, S] int64, "x_lens" [B], "y" [B, K, T] int64 (dataset-prepared: mask tokens, eog, empty-token delay pattern,
        audio_pad padding), "y_lens" [B]} -> {"loss", "top10acc", "top10acc_by_codebook", "effective_ntoken"} as the reference
        returns them (loss = sum_k mean-CE_k * ntokens_k * codebook_weight_k).  The loss masks are integer logic on the host
        (ssr.py:333-345); the transformer forward, the heads and the masked per-codebook cross entropy / top-10 accuracy run on
        the device (ssrb_lm_forward_loss), one utterance at its own length — the reference masks padded keys, so this is the
        same arithmetic."""
        x, x_lens, y, y_lens = batch["x"], batch["x_lens"], batch["y"], batch["y_lens"]
        if len(x) == 0:
            return None
        cfg, K = self.cfg, self.cfg.n_codebooks
        assert x.ndim == 2, x.shape
        assert x_lens.ndim == 1, x_lens.shape
        assert y.ndim == 3 and y.shape[1] == K, y.shape
        assert y_lens.ndim == 1, y_lens.shape
        if self._device is None:
            self.to("cuda")
        a = self.args
        predict_mask_token = bool(getattr(a, "predict_mask_token", 0))
        predict_all = bool(getattr(a, "predict_all", 0))
        cw = getattr(a, "codebook_weight", None)
        cw = [1.0] * K if cw is None else [float(v) for v in (eval(cw) if isinstance(cw, str) else cw)]
        xh = x.detach().to("cpu", torch.int64)
        yh = y.detach().to("cpu", torch.int64)
        Lmax, Tmax = int(x_lens.max()), int(y_lens.max())
        self._ensure_engine(2, Lmax + Tmax + 8, Lmax + Tmax + 16, 8)
        lib = _lib.load()
        sums = np.zeros((K, 4), dtype=np.float64)
        for b in range(xh.shape[0]):
            xl, yl = int(x_lens[b]), int(y_lens[b])
            xt = xh[b, :xl]
            yb = yh[b, :, :yl]
            if xt.numel() and (int(xt.min()) < 0 or int(xt.max()) >= cfg.n_text_tokens):
                raise IndexError("index out of range in self (phoneme id outside the embedding table)")
            flags = np.ascontiguousarray(seq.loss_flags(cfg, yb.numpy(), predict_mask_token, predict_all))     # ssr.py:333-345
            xt32 = xt.to(torch.int32).contiguous().numpy()
            yt32 = yb.to(torch.int32).contiguous().numpy()
            out = np.zeros((K, 4), dtype=np.float64)
            with torch.cuda.device(self._device):
                _lib.check(lib.ssrb_lm_forward_loss(self._h, C.c_void_p(xt32.ctypes.data), xl, C.c_void_p(yt32.ctypes.data), yl,
                                                    C.c_void_p(flags.ctypes.data), C.c_void_p(out.ctypes.data), _lib.stream_ptr()),
                           "forward_loss")
            sums += out
        dev = self._device
        loss_k = [sums[k, 0] / sums[k, 1] if sums[k, 1] > 0 else float("nan") for k in range(K)]       # mean over an empty set: nan, like F.cross_entropy
        acc_k = [sums[k, 2] / sums[k, 1] if sums[k, 1] > 0 else float("nan") for k in range(K)]
        nt = [int(sums[k, 3]) for k in range(K)]
        by_cb = [torch.tensor(acc_k[k] * nt[k], dtype=torch.float32, device=dev) for k in range(K)]
        return {"loss": torch.tensor(sum(loss_k[k] * nt[k] * cw[k] for k in range(K)), dtype=torch.float32, device=dev),
                "top10acc": torch.tensor(sum(acc_k[k] * nt[k] for k in range(K)), dtype=torch.float32, device=dev),
                "top10acc_by_codebook": by_cb,
                "effective_ntoken": torch.tensor(sum(nt), device=dev)}

    __call__ = forward

    def last_raw_logits(self) -> torch.Tensor:
        """[R, K, V] fp32 head outputs of the most recent iteration (before CFG / rules)."""
        K, V = self.cfg.n_codebooks, self.cfg.n_audio_tokens
        out = np.zeros((self._cap[0], K, V), dtype=np.float32)
        with torch.cuda.device(self._device):
            _lib.check(_lib.load().ssrb_lm_read_logits(self._h, _lib.stream_ptr(), C.c_void_p(out.ctypes.data)), "read_logits")
        return torch.from_numpy(out)

    def decode_path(self) -> int:
        """0 / 1 = per-GEMM chain (separate / folded LayerNorm), 2 = experimental per-layer kernel, 3 = persistent whole-iteration
        kernel for <= 16 rows (ssrb_lm_decode_path); -1 without an open batch."""
        return int(_lib.load().ssrb_lm_decode_path(self._h)) if self._h is not None else -1

    def step_bytes(self):
        wb, kb = C.c_double(0), C.c_double(0)
        with torch.cuda.device(self._device):
            _lib.check(_lib.load().ssrb_lm_step_bytes(self._h, _lib.stream_ptr(), C.byref(wb), C.byref(kb)), "step_bytes")
        return wb.value, kb.value
