"""Model hyper-parameters of the hot path.

SSRConfig mirrors the fields `SSR_Speech.__init__` reads from ``ckpt["config"]``
(reference models/ssr.py:113-179; 830M values from z_scripts/e830M.sh:21-26,38-41,55-68).
CodecConfig mirrors the WM-Encodec constructor kwargs
(audiocraft/config/model/encodec/default.yaml + encodec_large_nq4_s320.yaml).
"""
from __future__ import annotations

from argparse import Namespace
from dataclasses import dataclass, field
from typing import List, Tuple


@dataclass
class SSRConfig:
    d_model: int = 2048
    nhead: int = 16
    num_decoder_layers: int = 16
    audio_vocab_size: int = 2048
    n_codebooks: int = 4
    n_special: int = 5
    max_n_spans: int = 3
    text_vocab_size: int = 100
    text_pad_token: int = 100
    empty_token: int = 2048
    eog: int = 2049
    audio_pad_token: int = 2050
    eos: int = 2051
    sos: int = 2052
    mts: int = 2053

    @property
    def head_dim(self) -> int:
        return self.d_model // self.nhead

    @property
    def ffn_dim(self) -> int:  # models/ssr.py:163
        return 4 * self.d_model

    @property
    def head_hidden(self) -> int:  # models/ssr.py:177 (audio_vocab_size // 2)
        return self.audio_vocab_size // 2

    @property
    def n_audio_tokens(self) -> int:  # models/ssr.py:124
        return self.audio_vocab_size + self.n_special + self.max_n_spans

    @property
    def n_text_tokens(self) -> int:  # models/ssr.py:121
        return self.text_vocab_size + 1

    @staticmethod
    def from_args(args) -> "SSRConfig":
        """Accepts the argparse Namespace (or dict) stored in ``ckpt["config"]``."""
        d = dict(vars(args)) if isinstance(args, Namespace) else dict(args)
        V = d["audio_vocab_size"]
        if isinstance(V, str):  # models/ssr.py:118-119 evals a string
            V = int(eval(V))
        if not d.get("n_special", False):  # models/ssr.py:114-115
            d["n_special"] = 3
        cfg = SSRConfig(
            d_model=int(d["d_model"]), nhead=int(d["nhead"]), num_decoder_layers=int(d["num_decoder_layers"]),
            audio_vocab_size=int(V), n_codebooks=int(d["n_codebooks"]), n_special=int(d["n_special"]),
            max_n_spans=int(d["max_n_spans"]), text_vocab_size=int(d["text_vocab_size"]),
            text_pad_token=int(d["text_pad_token"]), empty_token=int(d["empty_token"]), eog=int(d["eog"]),
            audio_pad_token=int(d["audio_pad_token"]), eos=int(d.get("eos", -1)), sos=int(d["sos"]),
            mts=int(d["mts"]),
        )
        cfg.validate(audio_embedding_dim=int(d.get("audio_embedding_dim", cfg.d_model)))
        return cfg

    def validate(self, audio_embedding_dim=None):
        # the same assertions as models/ssr.py:122-130,195
        V = self.audio_vocab_size
        assert self.text_pad_token == self.text_vocab_size, (self.text_vocab_size, self.text_pad_token)
        assert V == self.empty_token, self.empty_token
        assert self.eog == V + 1, self.eog
        assert self.audio_pad_token == V + 2, self.audio_pad_token
        assert self.eos == V + 3, self.eos
        assert self.sos == V + 4, self.sos
        assert self.mts == V + 5, self.mts
        if audio_embedding_dim is not None:
            assert audio_embedding_dim == self.d_model, (audio_embedding_dim, self.d_model)
        assert self.d_model % self.nhead == 0
        assert self.head_dim == 128, "sm_100a attention kernels are specialised for head_dim=128"
        assert self.d_model % 128 == 0
        assert self.n_codebooks == 4, "sampling head is specialised for K=4 codebooks"
        assert 1 <= self.max_n_spans <= 3, "the engine's per-utterance span state holds at most 3 spans (SSRB_MAX_SPANS)"

    def to_namespace(self) -> Namespace:
        """Namespace accepted by the reference's SSR_Speech(args) (for fixtures / checkpoints)."""
        return Namespace(
            n_special=self.n_special, empty_token=self.empty_token, eog=self.eog,
            audio_pad_token=self.audio_pad_token, eos=self.eos, sos=self.sos, mts=self.mts,
            audio_vocab_size=str(self.audio_vocab_size), n_codebooks=self.n_codebooks,
            max_n_spans=self.max_n_spans, text_vocab_size=self.text_vocab_size,
            text_pad_token=self.text_pad_token, d_model=self.d_model, audio_embedding_dim=self.d_model,
            nhead=self.nhead, num_decoder_layers=self.num_decoder_layers,
            text_embedding_dropout=0.1, audio_embedding_dropout=0.0, text_positional_embedding_dropout=0.1,
            audio_positional_embedding_dropout=0.1, trm_dropout=0.1, shuffle_mask_embedding=0,
            predict_mask_token=1, predict_all=0, codebook_weight="[5,1,0.5,0.1]",
        )


def cfg_830m() -> SSRConfig:
    return SSRConfig()


def cfg_tiny(d_model=256, nhead=2, num_layers=2, audio_vocab_size=64) -> SSRConfig:
    V = audio_vocab_size
    return SSRConfig(d_model=d_model, nhead=nhead, num_decoder_layers=num_layers, audio_vocab_size=V,
                     empty_token=V, eog=V + 1, audio_pad_token=V + 2, eos=V + 3, sos=V + 4, mts=V + 5)


@dataclass
class CodecConfig:
    channels: int = 1
    dimension: int = 128
    n_filters: int = 64
    ratios: Tuple[int, ...] = (8, 5, 4, 2)   # decoder order; encoder uses them reversed (seanet.py:101)
    kernel_size: int = 7
    residual_kernel_size: int = 3
    last_kernel_size: int = 7
    compress: int = 2
    lstm: int = 2
    n_q: int = 4
    bins: int = 2048
    sample_rate: int = 16000

    @property
    def hop_length(self) -> int:
        h = 1
        for r in self.ratios:
            h *= r
        return h

    @property
    def frame_rate(self) -> int:
        return self.sample_rate // self.hop_length
