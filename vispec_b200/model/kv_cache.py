"""Static KV cache of the target LM -- mirror of vispec/model/kv_cache.py.

`KVCache` keeps the reference's interface (`.data`, `.current_length`, `.shape`, `.copy`, `.cat`), and
`initialize_past_key_values(model)` returns the same triple (per-layer [K, V] wrappers, the list of slabs,
the CPU length tensor).  The slab is the one the CUDA kernels write in place
([2*layers, 1, n_kv, max_pos, 128], kv_cache.py:118-126); lengths live on the device (vs_state_slot
VS_ST_CUR_LEN) with this CPU tensor kept as a coherent view for callers that read it
(vispec/evaluation/gen_baseline_answer_coco_caption.py:60-70, utils.py:541).
"""
from __future__ import annotations

import torch


class KVCache:
    """kv_cache.py:4-66."""

    def __init__(self, data: torch.Tensor, current_length: torch.Tensor):
        self.data = data
        self.current_length = current_length

    @property
    def shape(self):
        return (self.data.shape[0], self.data.shape[1], int(self.current_length.item()), self.data.shape[3])

    def copy(self, indices: torch.Tensor, prev_length: int, dim: int = 2):
        """kv_cache.py:38-50: gather `indices` along `dim` and store them contiguously at prev_length."""
        tgt = self.data.index_select(dim, indices.to(self.data.device))
        dst = self.data.narrow(dim, prev_length, tgt.shape[dim])
        dst.copy_(tgt, non_blocking=True)
        self.current_length.fill_(prev_length + tgt.shape[dim])

    def cat(self, tensor: torch.Tensor, dim: int = 2):
        """kv_cache.py:52-66.  The CUDA path appends inside the q/k/v projection's epilogue instead."""
        cur = int(self.current_length.item())
        dst = self.data.narrow(dim, cur, tensor.shape[dim])
        dst.copy_(tensor)
        self.current_length.add_(tensor.shape[dim])
        return torch.narrow(self.data, dim, 0, cur + tensor.shape[dim])


def initialize_past_key_values(model):
    """kv_cache.py:69-166.  `model` is the target LM wrapper (`SpecModel.base_model` or its `.language_model`)."""
    eng = getattr(model, "engine", None)
    if eng is None and hasattr(model, "language_model"):
        eng = getattr(model.language_model, "engine", None)
    if eng is None:
        raise TypeError("initialize_past_key_values expects a vispec_b200 target model")
    layers = eng.cfg.layers
    current_length_data = torch.zeros(layers * 2, dtype=torch.long, device="cpu")
    past_key_values = [[KVCache(eng.kv[2 * i + j], current_length_data[2 * i + j]) for j in range(2)]
                       for i in range(layers)]
    eng.set_len(0)
    return past_key_values, [eng.kv], current_length_data
