"""ssr-speech_b200 — B200-native (sm_100a) implementation of SSR-Speech's inference hot path:
WM-Encodec encode -> SSR_Speech.inference (prefill + AR decode + CFG + 4-codebook sampling head)
-> WM-Encodec decode / wmdecode, behind the reference's own Python API.

Import the pieces you need:
    from ssr_speech_b200.lm import SSR_Speech            # mirrors reference models/ssr.py::SSR_Speech
    from ssr_speech_b200.codec import AudioTokenizer     # mirrors reference data/tokenizer.py::AudioTokenizer
    from ssr_speech_b200.pipeline import inference_one_sample   # mirrors inference_scale.py
"""
__version__ = "0.1.0"
