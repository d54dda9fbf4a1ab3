"""The ViSpec draft model ("cnet") -- mirror of vispec/model/cnets_ours.py `Model`.

Same surface as the reference (`topK_genrate`, `reset_kv`, `init_tree`, `stable_kv`, `last_img_hidden`,
`total_tokens`, `depth`, `top_k`); every tensor op runs in the CUDA kernels of vispec_b200/csrc through
DraftEngine.  The draft's KV cache is written in place (no torch.cat realloc, cnets_ours.py:395-396) and the
tree build runs on the device (no .tolist() round trips, cnets_ours.py:1182-1204)."""
from __future__ import annotations

from typing import Optional

import torch

from .. import _lib as L
from ..engine import DraftConfig, DraftEngine


class Model:
    def __init__(self, config: DraftConfig, weights, lm_head_weight: torch.Tensor, device="cuda",
                 cache_rows: int = 4096, max_positions: int = 8192 + 64):
        self.config = config
        self.engine = DraftEngine(config, weights, lm_head_weight, device=device, cache_rows=cache_rows,
                                  max_positions=max_positions)
        self.total_tokens = config.total_tokens - 1       # cnets_ours.py:733
        self.depth = config.depth
        self.top_k = config.top_k
        self.num_q = config.num_q
        self.tree_mask = None
        self._has_stable = False

    # -- reference surface -----------------------------------------------------------------------------
    def init_tree(self):
        """cnets_ours.py:767-776 builds index helpers on the device; the kernels need none."""
        self.tree_mask = None

    def reset(self):
        self.tree_mask = None

    def reset_kv(self):
        """cnets_ours.py:1040-1041."""
        self._has_stable = False
        self.engine.reset_kv()

    @property
    def last_img_hidden(self) -> torch.Tensor:
        return self.engine.last_img_hidden

    @property
    def stable_kv(self):
        """((K [1,h,C,128], V, real_len),) like cnets_ours.py:1108, as views into the in-place cache."""
        if not self._has_stable:
            return None
        st = self.engine.state.tolist()
        c = st[L.ST_D_LEN]
        k = self.engine.k_cache[:, :c].unsqueeze(0)
        v = self.engine.v_cache[:, :c].unsqueeze(0)
        return ((k, v, st[L.ST_D_POS]),)

    @torch.no_grad()
    def topK_genrate(self, hidden_states, input_ids, head=None, logits_processor=None, inputs_embeds=None,
                     embed_weights=None, image_mask=None):
        """cnets_ours.py:1044-1238.  Returns (draft_tokens [1,T] i64, retrieve_indices [n_leaves,d] i64 (CPU),
        tree_mask [1,1,T,T] f32 (CPU), tree_position_ids [T] i64)."""
        eng = self.engine
        # T > 0 only changes the order of the leaf paths (cnets_ours.py:1215-1224): the tree build reads the flag
        import struct
        eng.state[L.ST_TEMP_BITS] = 0 if logits_processor is None else \
            struct.unpack("<i", struct.pack("<f", float(logits_processor.temperature)))[0]
        input_ids = input_ids.to(eng.device)
        sample_token = int(input_ids[0, -1].item())
        if not self._has_stable:
            if hidden_states.shape[0] != 1:
                raise NotImplementedError("batch size must be 1 (cnets_ours.py:901-902)")
            emb_tab = eng.embed
            if inputs_embeds is not None:                                     # :1066-1082
                n = inputs_embeds.shape[-2]
                new = torch.nn.functional.embedding(input_ids[:, n:], emb_tab)
                shifted = torch.cat((inputs_embeds[:, 1:].to(eng.device, eng.dtype), new), dim=-2)
            else:
                shifted = torch.nn.functional.embedding(input_ids[:, 1:], emb_tab)
            if shifted.shape[1] != hidden_states.shape[1]:
                raise ValueError("hidden_states and input_ids lengths do not match")
            eng.prefill(hidden_states, shifted, image_mask, sample_token)
            self._has_stable = True
        else:
            m = hidden_states.shape[1]
            ids = input_ids[0, 1:][-m:]
            eng.tree_from_accepted(hidden_states[0], ids, sample_token)
        return eng.tree_outputs()
