"""ctypes binding of libssr_b200.so (include/ssr_b200.h).  There is no CPU fallback: if the CUDA
library is missing or a call fails, this raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SSRB_LIB") or os.path.join(HERE, "libssr_b200.so")     # SSRB_LIB: A/B of two builds (tools/gpu_ab.sh)

SSRB_DTYPE_F32 = 0
SSRB_DTYPE_BF16 = 1
MAX_SPANS = 3
MAX_SILENCE = 8


class LMConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "d_model", "n_head", "n_layer", "ffn_dim", "n_codebooks", "n_audio_tokens", "n_text_tokens", "head_hidden",
        "empty_token", "eog", "eos", "sos", "mts", "max_n_spans",
        "max_rows", "max_seq", "max_prefill_tokens", "max_steps", "weight_dtype", "gemm_impl")]


class Sampling(C.Structure):
    _fields_ = [("top_k", C.c_int), ("top_p", C.c_float), ("temperature", C.c_float), ("stop_repetition", C.c_int),
                ("n_silence", C.c_int), ("silence_tokens", C.c_int * MAX_SILENCE), ("cfg_coef", C.c_float),
                ("cfg_stride", C.c_int), ("aug_text", C.c_int), ("seed", C.c_uint64)]


class LMBatch(C.Structure):
    _fields_ = [("n_utt", C.c_int), ("text", C.POINTER(C.c_int32)), ("text_stride", C.c_int),
                ("text_len", C.POINTER(C.c_int32)), ("prompt", C.POINTER(C.c_int32)), ("prompt_stride", C.c_int),
                ("prompt_len", C.POINTER(C.c_int32)), ("n_spans", C.POINTER(C.c_int32))]


class CodecConfigC(C.Structure):
    _fields_ = [("channels", C.c_int), ("dimension", C.c_int), ("n_filters", C.c_int), ("n_ratios", C.c_int),
                ("ratios", C.c_int * 8), ("kernel_size", C.c_int), ("residual_kernel_size", C.c_int),
                ("last_kernel_size", C.c_int), ("compress", C.c_int), ("lstm_layers", C.c_int), ("n_q", C.c_int),
                ("bins", C.c_int), ("max_batch_chunk", C.c_int), ("tensor_cores", C.c_int)]


EXPORTS = [
    "ssrb_last_error", "ssrb_version", "ssrb_launch_count",
    "ssrb_lm_create", "ssrb_lm_destroy", "ssrb_lm_load_tensor", "ssrb_lm_check_loaded", "ssrb_lm_begin",
    "ssrb_lm_decode", "ssrb_lm_poll", "ssrb_lm_admit", "ssrb_lm_poll_flags", "ssrb_lm_read_tokens", "ssrb_lm_read_logits", "ssrb_lm_decode_path", "ssrb_lm_teacher_forced", "ssrb_lm_forward_loss",
    "ssrb_lm_step_bytes", "ssrb_lm_profile_steps",
    "ssrb_codec_create", "ssrb_codec_destroy", "ssrb_codec_load_tensor", "ssrb_codec_check_loaded",
    "ssrb_codec_encode", "ssrb_codec_quantize", "ssrb_codec_decode", "ssrb_codec_wmdecode", "ssrb_codec_detect_watermark",
    "ssrb_op_gemm", "ssrb_op_gemm_ln", "ssrb_op_layer_chain", "ssrb_op_attn_decode", "ssrb_op_attn_prefill", "ssrb_debug_timeline", "ssrb_debug_mega_trace",
]

_lib = None


def load():
    """Loads the shared library (no CUDA call is made until an engine is created)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` (or ssr-speech_b200/build.py). "
            "ssr-speech_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, ip, i64p, fp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_float)
    lib.ssrb_last_error.restype = C.c_char_p
    lib.ssrb_launch_count.restype = C.c_uint64
    lib.ssrb_lm_create.argtypes = [C.POINTER(LMConfig), C.c_int, C.POINTER(vp)]
    lib.ssrb_lm_destroy.argtypes = [vp]
    lib.ssrb_lm_destroy.restype = None
    lib.ssrb_lm_load_tensor.argtypes = [vp, C.c_char_p, vp, i64p, C.c_int]
    lib.ssrb_lm_check_loaded.argtypes = [vp]
    lib.ssrb_lm_begin.argtypes = [vp, C.POINTER(LMBatch), C.POINTER(Sampling), vp, vp]
    lib.ssrb_lm_decode.argtypes = [vp, C.c_int, vp]
    lib.ssrb_lm_poll.argtypes = [vp, vp, ip, ip]
    lib.ssrb_lm_admit.argtypes = [vp, C.c_int, vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, vp]
    lib.ssrb_lm_poll_flags.argtypes = [vp, vp, vp, ip]
    lib.ssrb_lm_read_tokens.argtypes = [vp, vp, C.c_int, vp, C.c_int, ip, vp]
    lib.ssrb_lm_read_logits.argtypes = [vp, vp, vp]
    lib.ssrb_lm_decode_path.argtypes = [vp]
    lib.ssrb_lm_teacher_forced.argtypes = [vp, vp, C.c_int, vp, C.c_int, vp, vp]
    lib.ssrb_lm_forward_loss.argtypes = [vp, vp, C.c_int, vp, C.c_int, vp, vp, vp]
    lib.ssrb_lm_step_bytes.argtypes = [vp, vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.ssrb_lm_profile_steps.argtypes = [vp, C.c_int, vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.ssrb_codec_create.argtypes = [C.POINTER(CodecConfigC), C.c_int, C.POINTER(vp)]
    lib.ssrb_codec_destroy.argtypes = [vp]
    lib.ssrb_codec_destroy.restype = None
    lib.ssrb_codec_load_tensor.argtypes = [vp, C.c_char_p, vp, i64p, C.c_int]
    lib.ssrb_codec_check_loaded.argtypes = [vp]
    lib.ssrb_codec_encode.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp, vp]
    lib.ssrb_codec_quantize.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp]
    lib.ssrb_codec_decode.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp]
    lib.ssrb_codec_wmdecode.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, vp, vp, vp]
    lib.ssrb_codec_detect_watermark.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp]
    lib.ssrb_debug_timeline.argtypes = [vp, vp, C.c_uint]
    lib.ssrb_debug_mega_trace.argtypes = [vp, C.c_int, ip]
    lib.ssrb_op_gemm.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]
    lib.ssrb_op_gemm_ln.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, C.c_int, C.c_int, vp]
    lib.ssrb_op_layer_chain.argtypes = [vp] * 16 + [C.c_int] * 4 + [vp]
    lib.ssrb_op_attn_decode.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]
    lib.ssrb_op_attn_prefill.argtypes = [vp, vp, vp, vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, vp]
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().ssrb_last_error()
        raise RuntimeError(f"libssr_b200 {what} failed: {msg.decode() if msg else rc}")


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def load_state_dict_into(load_fn, handle, state_dict):
    """Feeds every floating-point tensor of a reference state_dict to ssrb_*_load_tensor."""
    import torch
    for name, t in state_dict.items():
        if not torch.is_tensor(t):
            continue
        t = t.detach().to("cpu", torch.float32).contiguous()
        if t.ndim == 0:
            t = t.reshape(1)
        shape = (C.c_int64 * t.ndim)(*t.shape)
        check(load_fn(handle, name.encode(), C.c_void_p(t.data_ptr()), shape, t.ndim), f"load_tensor({name})")
