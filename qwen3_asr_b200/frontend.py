"""GPU drop-in for the SDK processor's feature extractor.

The reference computes log-mel on the CPU inside the SDK processor
(``WhisperFeatureExtractor.__call__``, transformers feature_extraction_whisper.py:189-342, called with
``padding=True, truncation=False, return_attention_mask=True`` -- vllm transformers_utils/processors/qwen3_asr.py:114-130).
``B200FeatureExtractor`` keeps that call signature and return keys, and computes the features with the fused
CUDA log-mel kernel.  Semantics are per clip, standalone (reflect padding at both ends of every clip,
T = floor(N / 160)), i.e. what a single-clip request sees in the reference; frames
beyond a clip's length are zero and masked out.
"""

from __future__ import annotations

from typing import Sequence

import numpy as np
import torch


class B200FeatureExtractor:
    sampling_rate = 16000
    hop_length = 160
    n_fft = 400
    feature_size = 128

    def __init__(self, encoder, original=None):
        self.encoder = encoder    # B200AudioEncoder (owns the library handle)
        self.original = original  # the extractor this one replaced (server_hook.uninstall_frontend restores it)
        for attr in ("sampling_rate", "hop_length", "n_fft", "feature_size", "padding_value", "chunk_length", "n_samples",
                     "nb_max_frames", "return_attention_mask"):
            if original is not None and hasattr(original, attr):
                setattr(self, attr, getattr(original, attr))
        if (self.sampling_rate, self.hop_length, self.n_fft, self.feature_size) != (16000, 160, 400, 128):
            raise ValueError("the CUDA log-mel implements the Qwen3-ASR extractor only: 16 kHz, n_fft 400, hop 160, 128 mel bins")

    def __call__(self, raw_speech, sampling_rate: int | None = None, padding=True, truncation=False,
                 return_attention_mask=True, return_tensors=None, **kwargs):
        if sampling_rate is not None and sampling_rate != self.sampling_rate:
            raise ValueError(f"B200FeatureExtractor expects {self.sampling_rate} Hz audio, got {sampling_rate}")
        if truncation:
            raise ValueError("truncation is not supported (the reference calls the extractor with truncation=False)")
        if isinstance(raw_speech, (np.ndarray, torch.Tensor)) and raw_speech.ndim == 1:
            raw_speech = [raw_speech]
        clips: Sequence = [np.asarray(c, dtype=np.float32).reshape(-1) for c in raw_speech]
        if any(c.shape[0] <= self.n_fft // 2 for c in clips):
            raise ValueError("every clip needs more than 200 samples (torch.stft's reflect padding raises for them too)")
        mel, flens = self.encoder.logmel(clips)
        t_max = int(max(flens)) if len(flens) else 0
        feats = torch.zeros((len(clips), self.feature_size, t_max), dtype=torch.float32, device=mel.device)
        mask = torch.zeros((len(clips), t_max), dtype=torch.int32, device=mel.device)
        col = 0
        for i, t in enumerate(flens):
            t = int(t)
            feats[i, :, :t] = mel[:, col:col + t]
            mask[i, :t] = 1
            col += t
        out = {"input_features": feats}
        if return_attention_mask:
            out["attention_mask"] = mask
        if return_tensors == "np" or return_tensors is None:
            # WhisperFeatureExtractor hands back numpy unless asked for tensors (feature_extraction_whisper.py:339-342)
            out = {k: v.cpu().numpy() for k, v in out.items()}
        # same container as the extractor it replaces (a UserDict with .to(), feature_extraction_utils.BatchFeature); with
        # return_tensors="pt" the features stay on the GPU, so the SDK's later inputs.to(device) is a no-op
        from transformers.feature_extraction_utils import BatchFeature

        return BatchFeature(data=out)
