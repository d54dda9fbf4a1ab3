"""Draft-model config -- mirror of vispec/model/configs.py (`EConfig`) without the HF PretrainedConfig base.

Reads the same JSON files as the reference (vispec/train/*_config.json: hidden_size, intermediate_size,
num_attention_heads, vocab_size, max_position_embeddings, rms_norm_eps, rope_theta, bias)."""
from __future__ import annotations

import json
from dataclasses import dataclass


@dataclass
class EConfig:
    hidden_size: int = 4096
    intermediate_size: int = 11008
    num_hidden_layers: int = 1
    num_attention_heads: int = 32
    num_key_value_heads: int = 32
    vocab_size: int = 32064
    max_position_embeddings: int = 4096
    rms_norm_eps: float = 1e-5
    rope_theta: float = 10000.0
    bias: bool = True                      # spec_model_ours.py:59-64 defaults a missing key to True

    @classmethod
    def from_pretrained(cls, path: str) -> "EConfig":
        with open(path, "r") as f:
            raw = json.load(f)
        known = {k: raw[k] for k in cls.__dataclass_fields__ if k in raw}
        return cls(**known)
