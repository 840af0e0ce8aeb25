"""WM-Encodec on the B200: `WMEncodecModel`, `AudioTokenizer`, `tokenize_audio`.

Mirrors the reference's audio-side API for the inference path:
  * audiocraft/models/wmencodec.py::WMEncodecModel.{encode, decode, wmdecode, decode_latent}  (:324-386)
  * data/tokenizer.py::AudioTokenizer (:99-138) and tokenize_audio (:141-159)
  * checkpoint layout of audiocraft/solvers/wmcompression.py::model_from_checkpoint (:281-315):
    ``{'xp.cfg': cfg, 'best_state': {'model': state_dict}}``.
All conv / LSTM / RVQ arithmetic runs in libssr_b200.so (fp32, sm_100a); there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Optional, Tuple

import torch
import torch.nn.functional as F

from . import _lib
from .config import CodecConfig


class WMEncodecModel:
    def __init__(self, cfg: CodecConfig = CodecConfig(), max_batch_chunk: int = 8, precision: str = "fp32"):
        """precision: "fp32" = fp32 CUDA-core kernels everywhere (parity mode); "bf16" = decode / wmdecode on the
        tcgen05 tensor cores (bf16 operands, fp32 accumulate).  encode() is always fp32: its RVQ indices must be reproducible."""
        assert precision in ("fp32", "bf16"), precision
        self.precision = precision
        self.cfg = cfg
        self.sample_rate = cfg.sample_rate
        self.channels = cfg.channels
        self.frame_rate = cfg.frame_rate
        self.max_batch_chunk = max_batch_chunk
        self._sd = None
        self._device: Optional[torch.device] = None
        self._h = None

    # nn.Module-like surface
    def load_state_dict(self, state_dict, strict: bool = True):
        self._sd = {k: v.detach().to("cpu").contiguous() for k, v in state_dict.items() if torch.is_tensor(v)}
        if strict and "encoder.model.0.conv.conv.bias" not in self._sd:
            raise RuntimeError("Missing key(s) in state_dict: encoder.model.0.conv.conv.bias ...")
        self._destroy()
        return torch.nn.modules.module._IncompatibleKeys([], [])

    def state_dict(self):
        return dict(self._sd or {})

    def eval(self):
        return self

    def to(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("ssr_speech_b200 codec runs on CUDA devices only (no CPU fallback)")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        if device != self._device:
            self._destroy()
        self._device = device
        return self

    def _destroy(self):
        if self._h is not None:
            _lib.load().ssrb_codec_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    def _engine(self):
        if self._h is not None:
            return self._h
        if self._sd is None:
            raise RuntimeError("load_state_dict must be called first")
        if self._device is None:
            self.to("cuda")
        c = self.cfg
        ratios = (C.c_int * 8)(*(list(c.ratios) + [0] * (8 - len(c.ratios))))
        conf = _lib.CodecConfigC(channels=c.channels, dimension=c.dimension, n_filters=c.n_filters, n_ratios=len(c.ratios),
                                 ratios=ratios, kernel_size=c.kernel_size, residual_kernel_size=c.residual_kernel_size,
                                 last_kernel_size=c.last_kernel_size, compress=c.compress, lstm_layers=c.lstm, n_q=c.n_q,
                                 bins=c.bins, max_batch_chunk=self.max_batch_chunk,
                                 tensor_cores=1 if self.precision == "bf16" else 0)
        lib = _lib.load()
        h = C.c_void_p()
        with torch.cuda.device(self._device):
            _lib.check(lib.ssrb_codec_create(C.byref(conf), self._device.index, C.byref(h)), "ssrb_codec_create")
            self._h = h
            fsd = {k: v for k, v in self._sd.items() if v.is_floating_point()}
            _lib.load_state_dict_into(lib.ssrb_codec_load_tensor, h, fsd)
            _lib.check(lib.ssrb_codec_check_loaded(h), "ssrb_codec_check_loaded")
        return self._h

    # ---- reference API -----------------------------------------------------------------------------------
    @torch.no_grad()
    def encode(self, x: torch.Tensor) -> Tuple[torch.Tensor, Optional[torch.Tensor], torch.Tensor]:
        """x [B,1,T] float -> (codes [B,K,T/hop] int64, scale None (renormalize=False), emb [B,D,T/hop])."""
        assert x.dim() == 3
        h = self._engine()
        x = x.to(self._device, torch.float32).contiguous()
        B, Cc, T = x.shape
        assert Cc == self.cfg.channels
        Tf = T
        for r in reversed(self.cfg.ratios):        # ceil chain of the strided convs (= T // hop when hop divides T)
            Tf = -(-Tf // r)
        codes = torch.empty(B, self.cfg.n_q, Tf, dtype=torch.int64, device=self._device)
        emb = torch.empty(B, self.cfg.dimension, Tf, dtype=torch.float32, device=self._device)
        with torch.cuda.device(self._device):
            _lib.check(_lib.load().ssrb_codec_encode(h, C.c_void_p(x.data_ptr()), B, T, C.c_void_p(codes.data_ptr()),
                                                     C.c_void_p(emb.data_ptr()), _lib.stream_ptr()), "codec_encode")
        return codes, None, emb

    @torch.no_grad()
    def quantize(self, emb: torch.Tensor) -> torch.Tensor:
        """RVQ encode of given latents [B,D,Tf] -> codes [B,K,Tf] (quantization/core_vq.py:382-392)."""
        h = self._engine()
        emb = emb.to(self._device, torch.float32).contiguous()
        B, D, Tf = emb.shape
        codes = torch.empty(B, self.cfg.n_q, Tf, dtype=torch.int64, device=self._device)
        with torch.cuda.device(self._device):
            _lib.check(_lib.load().ssrb_codec_quantize(h, C.c_void_p(emb.data_ptr()), B, Tf, C.c_void_p(codes.data_ptr()),
                                                       _lib.stream_ptr()), "codec_quantize")
        return codes

    @torch.no_grad()
    def decode(self, codes: torch.Tensor, scale: Optional[torch.Tensor] = None) -> torch.Tensor:
        assert scale is None, "renormalize=False for the SSR-Speech codec (encodec_audiogen_16khz.yaml:9-10)"
        h = self._engine()
        codes = codes.to(self._device, torch.int64).contiguous()
        B, K, Tf = codes.shape
        wav = torch.empty(B, 1, Tf * self.cfg.hop_length, dtype=torch.float32, device=self._device)
        with torch.cuda.device(self._device):
            _lib.check(_lib.load().ssrb_codec_decode(h, C.c_void_p(codes.data_ptr()), B, Tf, C.c_void_p(wav.data_ptr()),
                                                     _lib.stream_ptr()), "codec_decode")
        return wav

    @torch.no_grad()
    def wmdecode(self, codes: torch.Tensor, labels: torch.Tensor, wavform: torch.Tensor,
                 scale: Optional[torch.Tensor] = None, return_marks: bool = True):
        assert scale is None
        h = self._engine()
        codes = codes.to(self._device, torch.int64).contiguous()
        labels = labels.to(self._device, torch.int64).contiguous()
        wavform = wavform.to(self._device, torch.float32).contiguous()
        B, K, Tf = codes.shape
        T = Tf * self.cfg.hop_length
        assert labels.shape == (B, Tf), labels.shape
        assert wavform.shape == (B, 1, T), wavform.shape
        out = torch.empty(B, 1, T, dtype=torch.float32, device=self._device)
        marks = torch.empty(B, Tf, 2, dtype=torch.float32, device=self._device) if return_marks else None
        with torch.cuda.device(self._device):
            _lib.check(_lib.load().ssrb_codec_wmdecode(
                h, C.c_void_p(codes.data_ptr()), C.c_void_p(labels.data_ptr()), C.c_void_p(wavform.data_ptr()), B, Tf,
                C.c_void_p(out.data_ptr()), C.c_void_p(marks.data_ptr()) if marks is not None else None,
                _lib.stream_ptr()), "codec_wmdecode")
        return out, marks


def _detect_watermark(self, x: torch.Tensor, reference_axis_bug: bool = True):
    """WMEncodecModel.detect_watermark (wmencodec.py:377-382).  Returns (marks, logits [B,Tf,2]).
    reference_axis_bug=True reproduces the reference exactly: argmax over dim=-1 of logits laid out [B,2,Tf] (i.e. over TIME,
    shape [B,2]); False returns the per-frame class decision [B,Tf] the code evidently intended."""
    assert x.dim() == 3
    h = self._engine()
    x = x.to(self._device, torch.float32).contiguous()
    B, _, T = x.shape
    Tf = T // self.cfg.hop_length
    logits = torch.empty(B, Tf, 2, dtype=torch.float32, device=self._device)
    with torch.cuda.device(self._device):
        _lib.check(_lib.load().ssrb_codec_detect_watermark(h, C.c_void_p(x.data_ptr()), B, T, C.c_void_p(logits.data_ptr()),
                                                           _lib.stream_ptr()), "codec_detect_watermark")
    marks = torch.argmax(logits.transpose(1, 2), dim=-1) if reference_axis_bug else torch.argmax(logits, dim=-1)
    return marks, logits


WMEncodecModel.detect_watermark = _detect_watermark


def _cfg_from_xp(xp_cfg) -> CodecConfig:
    """Reads the resolved Hydra cfg stored in the checkpoint (wmcompression.py:302-304; consumed by
    models/builders.py:68-113: cfg.seanet -> SEANet kwargs, cfg.rvq -> quantizer, cfg.sample_rate / cfg.channels).
    Strict: a missing section / key or a value these kernels do not implement raises instead of silently decoding with the
    default geometry (other ratios or bins would produce garbage without a word)."""
    def get(o, k, where):
        try:
            if isinstance(o, dict) or hasattr(o, "keys"):
                if k in o:
                    return o[k]
            elif hasattr(o, k):
                return getattr(o, k)
        except Exception as e:
            raise ValueError(f"malformed xp.cfg: cannot read {where}.{k}: {e!r}") from e
        raise ValueError(f"malformed xp.cfg: {where}.{k} is missing")

    def opt(o, k, default):
        try:
            return get(o, k, "")
        except ValueError:
            return default
    if xp_cfg is None:
        raise ValueError("malformed checkpoint: 'xp.cfg' is None")
    se, rv = get(xp_cfg, "seanet", "xp.cfg"), get(xp_cfg, "rvq", "xp.cfg")
    try:
        cfg = CodecConfig(
            channels=int(get(xp_cfg, "channels", "xp.cfg")), dimension=int(get(se, "dimension", "xp.cfg.seanet")),
            n_filters=int(get(se, "n_filters", "xp.cfg.seanet")),
            ratios=tuple(int(r) for r in get(se, "ratios", "xp.cfg.seanet")),
            kernel_size=int(get(se, "kernel_size", "xp.cfg.seanet")),
            residual_kernel_size=int(get(se, "residual_kernel_size", "xp.cfg.seanet")),
            last_kernel_size=int(get(se, "last_kernel_size", "xp.cfg.seanet")), compress=int(get(se, "compress", "xp.cfg.seanet")),
            lstm=int(get(se, "lstm", "xp.cfg.seanet")), n_q=int(get(rv, "n_q", "xp.cfg.rvq")), bins=int(get(rv, "bins", "xp.cfg.rvq")),
            sample_rate=int(get(xp_cfg, "sample_rate", "xp.cfg")))
    except (TypeError, AttributeError) as e:
        raise ValueError(f"malformed xp.cfg: {e!r}") from e
    # what the kernels are specialised for; present-and-different is an error, absent means default
    fixed = {"n_residual_layers": 1, "activation": "ELU", "norm": "weight_norm", "dilation_base": 2, "pad_mode": "constant",
             "true_skip": True, "causal": False, "disable_norm_outer_blocks": 0}
    for k, want in fixed.items():
        v = opt(se, k, want)
        if v != want:
            raise ValueError(f"xp.cfg.seanet.{k}={v!r} is not supported by the sm_100a codec kernels (need {want!r})")
    enc = opt(xp_cfg, "encodec", None)
    if enc is not None and bool(opt(enc, "renormalize", False)):
        raise ValueError("xp.cfg.encodec.renormalize=True is not supported (the SSR-Speech codec uses renormalize=False)")
    return cfg


class AudioTokenizer:
    """EnCodec audio tokenizer — same surface as reference data/tokenizer.py:99-138."""

    def __init__(self, device: Any = None, signature=None, model: Optional[WMEncodecModel] = None):
        if model is None:
            state = torch.load(signature, map_location="cpu", weights_only=False)
            assert state is not None and "xp.cfg" in state, f"Could not load compression model from ckpt: {signature}"
            assert "best_state" in state and state["best_state"] != {}
            model = WMEncodecModel(_cfg_from_xp(state["xp.cfg"]))
            model.load_state_dict(state["best_state"]["model"])
        self.sample_rate = model.sample_rate
        self.channels = model.channels
        if not device:
            if not torch.cuda.is_available():
                raise RuntimeError("ssr_speech_b200 needs a CUDA device (no CPU fallback)")
            device = torch.device("cuda:0")
        self._device = torch.device(device)
        self.codec = model.to(self._device)

    @property
    def device(self):
        return self._device

    def encode(self, wav: torch.Tensor):
        return self.codec.encode(wav.to(self.device))

    def decode(self, frames: torch.Tensor, scale: Optional[torch.Tensor] = None) -> torch.Tensor:
        return self.codec.decode(frames, scale)

    def wmdecode(self, frames: torch.Tensor, marks: torch.Tensor, wav: torch.Tensor, scale: Optional[torch.Tensor] = None):
        out, _ = self.codec.wmdecode(frames.to(self.device), marks.to(self.device), wav.to(self.device), scale,
                                     return_marks=False)
        return out

    def detect_watermark(self, wav: torch.Tensor):
        marks, _ = self.codec.detect_watermark(wav.to(self.device))
        return marks


def load_wav(path: str, offset: int = -1, num_frames: int = -1):
    """torchaudio.load replacement (torchaudio's backends are absent in this image): float32 [C, T], sr."""
    try:
        import torchaudio
        if offset != -1 and num_frames != -1:
            return torchaudio.load(path, frame_offset=offset, num_frames=num_frames)
        return torchaudio.load(path)
    except Exception:
        from scipy.io import wavfile
        import numpy as np
        sr, data = wavfile.read(path)
        if data.dtype == np.int16:
            data = data.astype(np.float32) / 32768.0
        elif data.dtype == np.int32:
            data = data.astype(np.float32) / 2147483648.0
        data = torch.from_numpy(np.asarray(data, dtype=np.float32))
        data = data[None] if data.ndim == 1 else data.T
        if offset != -1 and num_frames != -1:
            data = data[:, offset:offset + num_frames]
        return data.contiguous(), sr


def pad_to_multiple(wav: torch.Tensor, multiple: int = 320) -> torch.Tensor:
    """data/tokenizer.py:148-151."""
    pad = (multiple - (wav.shape[-1] % multiple)) % multiple
    return F.pad(wav, (0, pad), "constant", 0) if pad > 0 else wav


def convert_audio(wav: torch.Tensor, sr: int, target_sr: int, target_channels: int) -> torch.Tensor:
    """data/tokenizer.py:87-97 (resampling is a no-op at 16 kHz; other rates need torchaudio)."""
    assert wav.shape[0] in [1, 2], "Audio must be mono or stereo."
    if target_channels == 1:
        wav = wav.mean(0, keepdim=True)
    elif target_channels == 2:
        wav = wav.expand(target_channels, wav.shape[-1])
    if sr != target_sr:
        import torchaudio
        wav = torchaudio.transforms.Resample(sr, target_sr)(wav)
    return wav


def tokenize_audio(tokenizer: AudioTokenizer, audio_path, offset=-1, num_frames=-1, multiple=320):
    """data/tokenizer.py:141-159; `audio_path` may also be a float tensor [C, T] at the codec rate."""
    if torch.is_tensor(audio_path):
        wav, sr = audio_path, tokenizer.sample_rate
    else:
        wav, sr = load_wav(audio_path, offset, num_frames)
    wav = pad_to_multiple(wav, multiple)
    wav = convert_audio(wav, sr, tokenizer.sample_rate, tokenizer.channels)
    wav = wav.unsqueeze(0)
    with torch.no_grad():
        encoded_frames, scale, emb = tokenizer.encode(wav)
    return encoded_frames, scale, emb
