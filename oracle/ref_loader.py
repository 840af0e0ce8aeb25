"""TEST INFRASTRUCTURE — loader for the *unmodified* reference (container only).

Imports the reference's own Python modules from ``/root/reference`` by path, with
the three stubs SURVEY.md §8(c)/Appendix E describes (torchmetrics, flashy,
synthetic parent packages for audiocraft so that its heavyweight ``__init__``
chain — xformers/omegaconf/dora — is never executed).  No reference file is
modified or copied.

Used only by ``oracle/gen_golden.py`` (fixture generation) and by
``bench.py --impl reference`` / ``cpu_baseline`` when ``/root/reference`` exists
(it does not exist on the GPU box; callers must fall back to the oracle port).
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from argparse import Namespace

def _find_reference_root() -> str:
    """SSRB_REFERENCE_ROOT, then /root/reference (build container), then <repo>/baseline/_ref (a checkout placed beside the repo,
    git-ignored).  The reference is a script tree, not an installable package, so there is nothing to pip-install."""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for cand in (os.environ.get("SSRB_REFERENCE_ROOT"), "/root/reference", os.path.join(here, "baseline", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "models", "ssr.py")):
            return cand
    return os.environ.get("SSRB_REFERENCE_ROOT") or "/root/reference"


REF_ROOT = _find_reference_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "models", "ssr.py"))


def _stub_torchmetrics():
    if "torchmetrics" in sys.modules:
        return
    import torch.nn as nn

    class MulticlassAccuracy(nn.Module):  # train-time metric only (models/ssr.py:12,181-189)
        """Functional stand-in for torchmetrics.classification.MulticlassAccuracy as SSR_Speech.forward uses it
        (top_k=10, average="micro", multidim_average="global", ignore_index=None on [N, C] logits and [N] targets):
        the fraction of samples whose target is among the top_k scores."""

        def __init__(self, num_classes=None, top_k=1, average="micro", **k):
            super().__init__()
            assert average == "micro"
            self.top_k = int(top_k)

        def forward(self, preds, target):
            top = preds.topk(self.top_k, dim=-1).indices
            return (top == target[:, None]).any(-1).float().mean()

    tm = types.ModuleType("torchmetrics")
    tmc = types.ModuleType("torchmetrics.classification")
    tmc.MulticlassAccuracy = MulticlassAccuracy
    tm.classification = tmc
    sys.modules["torchmetrics"] = tm
    sys.modules["torchmetrics.classification"] = tmc


def load_reference_ssr():
    """Returns the reference ``models.ssr`` module (SSR_Speech, topk_sampling, ...)."""
    assert reference_available(), f"reference not found under {REF_ROOT}"
    _stub_torchmetrics()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    return importlib.import_module("models.ssr")


def ssr_args(d_model=2048, nhead=16, num_layers=16, audio_vocab_size=2048, n_codebooks=4,
             text_vocab_size=100, max_n_spans=3) -> Namespace:
    """Namespace with the fields read at models/ssr.py:113-179 (values: z_scripts/e830M.sh)."""
    V = int(audio_vocab_size)
    return Namespace(
        n_special=5, empty_token=V, eog=V + 1, audio_pad_token=V + 2, eos=V + 3, sos=V + 4, mts=V + 5,
        audio_vocab_size=str(V), n_codebooks=n_codebooks, max_n_spans=max_n_spans,
        text_vocab_size=text_vocab_size, text_pad_token=text_vocab_size,
        d_model=d_model, audio_embedding_dim=d_model, nhead=nhead, num_decoder_layers=num_layers,
        text_embedding_dropout=0.1, audio_embedding_dropout=0.0,
        text_positional_embedding_dropout=0.1, audio_positional_embedding_dropout=0.1, trm_dropout=0.1,
        shuffle_mask_embedding=0, predict_mask_token=1, predict_all=0, codebook_weight="[5,1,0.5,0.1]",
    )


def load_reference_codec_modules():
    """Returns (seanet, quantization, wmencodec) reference modules loaded by path."""
    assert reference_available(), f"reference not found under {REF_ROOT}"
    R = os.path.join(REF_ROOT, "audiocraft", "audiocraft")
    if "flashy" not in sys.modules:
        fl = types.ModuleType("flashy")
        fl.distrib = types.SimpleNamespace(broadcast_tensors=lambda *a, **k: None)
        sys.modules["flashy"] = fl
    for name, path in (("ac", R), ("ac.modules", R + "/modules"), ("ac.models", R + "/models")):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [path]
            sys.modules[name] = m
    seanet = importlib.import_module("ac.modules.seanet")
    qt = importlib.import_module("ac.quantization")
    sys.modules["ac"].quantization = qt
    wm = importlib.import_module("ac.models.wmencodec")
    return seanet, qt, wm


CODEC_KW = dict(channels=1, dimension=128, n_filters=64, n_residual_layers=1, ratios=[8, 5, 4, 2],
                activation="ELU", activation_params={"alpha": 1.0}, norm="weight_norm", norm_params={},
                kernel_size=7, residual_kernel_size=3, last_kernel_size=7, dilation_base=2,
                pad_mode="constant", true_skip=True, compress=2, lstm=2, disable_norm_outer_blocks=0,
                causal=False)


def build_reference_codec(n_filters=64, dimension=128, bins=2048, n_q=4, ratios=(8, 5, 4, 2)):
    """WMEncodecModel built with the kwargs of SURVEY Appendix A.2
    (audiocraft/config/model/encodec/default.yaml + encodec_large_nq4_s320.yaml; builders.py:68-113)."""
    seanet, qt, wm = load_reference_codec_modules()
    kw = dict(CODEC_KW, n_filters=n_filters, dimension=dimension, ratios=list(ratios))
    enc = seanet.SEANetEncoder(**kw)
    dkw = dict(kw, trim_right_ratio=1.0, final_activation=None)
    dec = seanet.SEANetDecoder(**dkw)
    wmdec = seanet.WMSEANetDecoder(**dkw)
    q = qt.ResidualVectorQuantizer(dimension=dimension, n_q=n_q, q_dropout=False, bins=bins, decay=0.99,
                                   kmeans_init=True, kmeans_iters=50, threshold_ema_dead_code=2,
                                   orthogonal_reg_weight=0.0, orthogonal_reg_active_codes_only=False,
                                   orthogonal_reg_max_codes=None)
    hop = 1
    for r in ratios:
        hop *= r
    model = wm.WMEncodecModel(enc, dec, wmdec, q, frame_rate=16000 // hop, sample_rate=16000, channels=1,
                              causal=False, renormalize=False)
    return model.eval()
