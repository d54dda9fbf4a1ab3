"""Spec-decode cycle functions -- mirror of vispec/model/utils.py (same names, arguments and return tuples).

These are the host-visible, one-stage-at-a-time entry points (`initialize_tree`, `tree_decoding`,
`evaluate_posterior`, `update_inference_inputs`, `reset_tree_mode`, `prepare_logits_processor`, `Timer`).
Each stage runs in the CUDA kernels behind the C ABI and, like the reference, hands host-visible tensors
back to the caller (which costs a sync per stage).  `SpecModel.specgenerate` does not use them: it runs
the fused, sync-free `vs_spec_cycle`.  Both routes execute the same kernels.
"""
from __future__ import annotations

import ctypes as C
import time
from typing import List

import torch

from .. import _lib as L
from ..engine import _stream


class Timer:
    """utils.py:25-36."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        torch.cuda.synchronize()
        self.start = time.perf_counter()

    def __exit__(self, exc_type, exc_value, traceback):
        torch.cuda.synchronize()
        print(f"{self.name} took {time.perf_counter() - self.start} seconds")


class TemperatureProcessor:
    """What utils.py:39-55 builds for the sampling runs: a LogitsProcessorList holding a TemperatureLogitsWarper and,
    for top_k > 0, a TopKLogitsWarper (every logit below the k-th largest of its row becomes -inf; ties with the k-th
    value stay).  Callable like it (`proc(None, logits)`); the CUDA path reads `.temperature` and `.top_k`."""

    def __init__(self, temperature: float, top_k: int = 0):
        self.temperature = float(temperature)
        self.top_k = int(top_k)

    def __call__(self, input_ids, scores):
        scores = scores / self.temperature
        if self.top_k > 0:
            k = min(self.top_k, scores.shape[-1])
            kth = torch.topk(scores, k, dim=-1).values[..., -1:]
            scores = scores.masked_fill(scores < kth, float("-inf"))
        return scores


def prepare_logits_processor(temperature: float = 0.0, repetition_penalty: float = 0.0, top_p: float = 0.0,
                             top_k: int = 0):
    """utils.py:39-55.  The reference returns an (empty) processor list for temperature <= 1e-5 and callers only build
    one for temperature > 1e-5 (spec_model_ours.py:272-277); here: None for greedy, a TemperatureProcessor (temperature
    + optional top-k) for sampling.  The nucleus (top-p) and repetition-penalty warpers are not on the B200 path: no
    evaluation script sets them, and the reference's own nucleus filter raises on the 3-D tree logits of
    evaluate_posterior (tests/golden/warpers_sampling.pt records it)."""
    if temperature <= 1e-5:
        return None
    if repetition_penalty > 1.0 or 1e-8 <= top_p < 1.0:
        raise NotImplementedError("temperature and top_k sampling are implemented on the B200 path "
                                  "(top_p / repetition_penalty warpers are not)")
    return TemperatureProcessor(temperature, int(top_k) if top_k and top_k > 0 else 0)


def _temp_bits(logits_processor) -> int:
    import struct
    return 0 if logits_processor is None else struct.unpack("<i", struct.pack("<f", logits_processor.temperature))[0]


TOPK = 10          # sibling slots per group of the static-tree candidate buffer (utils.py:14)


def generate_tree_buffers(tree_choices, device="cuda"):
    """Static draft-tree buffers of the EAGLE-1 / Medusa style decode that the baseline evaluation scripts still build
    and cache on the model (`gen_baseline_answer_coco_caption.py:48-70`; reference: `utils.py:90-218`).  Host-side integer
    bookkeeping only -- the dynamic tree of the draft-and-verify path is built on the device (`vs_tree_build`).

    `tree_choices`: list of paths, each a list of child ranks from the root, e.g. [[0], [1], [0, 0]].  Returns tensors on
    `device`: `tree_attn_mask` [1,1,n,n] (node sees itself, the root and its ancestors), `tree_indices` [n] (slot of each
    node in the flat candidate buffer: rank + TOPK * sibling-group number + 1), `tree_position_ids` [n] (depth),
    `retrieve_indices` [leaves, depth+1] (root-to-leaf node indices, -1 padded, rows in ascending order with the padding
    last); node 0 is the root, nodes follow in (depth, path) order."""
    order = sorted((tuple(c) for c in tree_choices), key=lambda c: (len(c), c))
    node = {c: i + 1 for i, c in enumerate(order)}
    n = len(order) + 1
    mask = torch.zeros(n, n)
    mask[:, 0] = 1
    pos = torch.zeros(n, dtype=torch.long)
    slot = torch.zeros(n, dtype=torch.long)
    group, prev = -1, None
    interior = set()
    for c in order:
        i = node[c]
        mask[i, i] = 1
        for k in range(1, len(c)):
            mask[i, node[c[:k]]] = 1
            interior.add(c[:k])
        pos[i] = len(c)
        if (len(c), c[:-1]) != prev:              # a new run of siblings (same depth, same parent)
            group, prev = group + 1, (len(c), c[:-1])
        slot[i] = c[-1] + TOPK * group + 1
    leaves = [c for c in order if c not in interior]
    width = max((len(c) for c in leaves), default=0) + 1
    rows = [[0] + [node[c[:k]] for k in range(1, len(c) + 1)] for c in leaves]
    rows = [r + [-1] * (width - len(r)) for r in rows]
    big = n + 5
    rows.sort(key=lambda r: [x if x >= 0 else big for x in r])
    out = {"tree_attn_mask": mask[None, None], "tree_indices": slot, "tree_position_ids": pos,
           "retrieve_indices": torch.tensor(rows, dtype=torch.long).reshape(len(rows), width)}
    return {k: v.clone().to(device) for k, v in out.items()}


def reset_tree_mode(model):
    """utils.py:330-339."""
    lm = model.base_model
    lm = getattr(lm, "language_model", lm)
    lm.model.tree_mask = None
    lm.model.tree_mode = None


def reset_past_key_values(passed_key_values: List[torch.Tensor]) -> List[torch.Tensor]:
    """utils.py:342-360."""
    for i in range(len(passed_key_values)):
        for j in range(2):
            passed_key_values[i][j].current_length.fill_(0)
    return passed_key_values


def _lm_head(model):
    lm = model.base_model
    return getattr(lm, "language_model", lm).lm_head


def initialize_tree(input_ids, model, past_key_values, logits_processor, inputs_embeds=None, embed_weights=None,
                    image_mask=None, **kwargs):
    """utils.py:266-327: target prefill, first token, first draft tree."""
    outputs, orig, hidden_states = model(input_ids, past_key_values=past_key_values, output_orig=True,
                                         inputs_embeds=inputs_embeds, position_ids=kwargs.get("position_ids"))
    if logits_processor is not None:                                        # utils.py:284-288
        # stage-by-stage route: a fresh Philox key per generation, drawn from torch's generator (torch.manual_seed makes
        # the run reproducible, two generations do not share a uniform stream); specgenerate reseeds on its own
        if hasattr(model, "sampling_setup") and not getattr(model, "_vs_seeded", False):
            model.sampling_setup(logits_processor, reseed=True)
        probabilities = torch.nn.functional.softmax(logits_processor(None, orig[:, -1]), dim=1)
        token = torch.multinomial(probabilities, 1)
    else:
        token = torch.argmax(orig[:, -1])[None, None]
    input_ids = torch.cat((input_ids, token.to(input_ids.device)), dim=1)
    draft_tokens, retrieve_indices, tree_mask, tree_position_ids = model.spec_layer.topK_genrate(
        hidden_states, input_ids, _lm_head(model), logits_processor, inputs_embeds=inputs_embeds,
        image_mask=image_mask)
    return draft_tokens, retrieve_indices, tree_mask, tree_position_ids, orig, hidden_states, token


def tree_decoding(model, tree_candidates, past_key_values, tree_position_ids, input_ids, retrieve_indices):
    """utils.py:389-412: verify the tree with the target (tree mask taken from `base_model.model.tree_mask`)."""
    position_ids = tree_position_ids + input_ids.shape[1]
    if getattr(model.base_model, "rope_deltas", 0):                        # utils.py:398-402 (Qwen2.5-VL)
        position_ids = position_ids + model.base_model.rope_deltas
    outputs, tree_logits, hidden_state = model(tree_candidates, output_orig=True, past_key_values=past_key_values,
                                               position_ids=position_ids)
    logits = tree_logits[0, retrieve_indices.to(tree_logits.device)]
    logits._vs_model = model                 # lets evaluate_posterior use the device-side arg-max of this verify
    logits._vs_prev_len = int(input_ids.shape[1])     # kv length before the tree rows were appended
    return logits, hidden_state, outputs


def evaluate_posterior(logits: torch.Tensor, candidates: torch.Tensor, logits_processor):
    """utils.py:415-493: greedy branch (:438-451) through vs_accept_greedy; sampling branch (:453-493) through the
    accept-sampling phase, which also draws the next root token (utils.py:550-552), so `sample_p` is returned for its
    shape only and carries that token."""
    model = getattr(logits, "_vs_model", None)
    if model is None:
        raise TypeError("evaluate_posterior expects the logits returned by vispec_b200's tree_decoding")
    eng_t, eng_d = model.base_model_engine, model.spec_layer.engine
    # the verify call left the device-side kv length at prev + tree rows (as the reference's KVCache does); the accept
    # kernels count from the length BEFORE the tree: PREV_LEN = prev, CUR_LEN = prev + accepted + 1
    prev_len = getattr(logits, "_vs_prev_len", None)
    if prev_len is not None:
        eng_t.set_len(prev_len)
    if logits_processor is not None:
        model.sampling_setup(logits_processor)
        model.accept_sampling_program(logits_processor.temperature, getattr(logits_processor, "top_k", 0)).run()
    else:
        L.check(eng_d.lib.vs_accept_greedy(C.byref(eng_d.tb), L.ptr(eng_t.state), L.ptr(eng_t.amax), L.ptr(eng_d.out_rec),
                                           _stream()), "vs_accept_greedy")
    st = eng_t.state.tolist()
    accept_length, best = st[L.ST_ACCEPT], st[L.ST_BEST]
    best_candidate = torch.tensor(best, dtype=torch.long, device=candidates.device)
    sample_p = logits[best, accept_length]
    sample_p._vs_next_token = st[L.ST_NEXT_TOKEN]
    return best_candidate, torch.tensor(accept_length, device=logits.device), sample_p


@torch.no_grad()
def update_inference_inputs(input_ids, candidates, best_candidate, accept_length, retrieve_indices, logits_processor,
                            new_token, past_key_values_data_list, current_length_data, model, hidden_state_new,
                            sample_p):
    """utils.py:497-593: append the accepted tokens, splice the KV cache, build the next draft tree."""
    eng_t = model.base_model_engine
    prev_input_len = input_ids.shape[1]
    a = int(accept_length)
    best = int(best_candidate)
    sel = retrieve_indices[best, :a + 1]
    input_ids = torch.cat([input_ids, candidates[None, best, :a + 1].to(input_ids.device)], dim=-1)
    sel_host = (C.c_int32 * L.VS_MAX_ACCEPT)(*[int(x) for x in sel.tolist()])
    for slab in past_key_values_data_list:                                   # utils.py:529-541
        n_slabs, _, n_kv, max_pos, _ = slab.shape
        L.check(eng_t.lib.vs_kv_splice(L.ptr(slab), n_slabs, n_kv, max_pos, None, None, prev_input_len, a + 1,
                                       sel_host, _stream()), "vs_kv_splice")
    current_length_data.fill_(prev_input_len + a + 1)
    eng_t.set_len(prev_input_len + a + 1)
    accept_hidden_state_new = hidden_state_new[:, sel.to(hidden_state_new.device)]      # utils.py:543-546
    tok = getattr(sample_p, "_vs_next_token", None)
    if tok is None:
        raise TypeError("update_inference_inputs expects the sample_p returned by vispec_b200's evaluate_posterior")
    token = torch.tensor(tok, device=input_ids.device)[None, None]
    draft_tokens, retrieve_indices, tree_mask, tree_position_ids = model.spec_layer.topK_genrate(
        accept_hidden_state_new, input_ids=torch.cat((input_ids, token.to(input_ids.device)), dim=1),
        head=_lm_head(model), logits_processor=logits_processor)
    new_token += a + 1
    return input_ids, draft_tokens, retrieve_indices, tree_mask, tree_position_ids, new_token, None, token


def qwen_rope_index(input_ids: torch.Tensor, image_grid_thw=None, image_token_id: int = 151655,
                    spatial_merge_size: int = 2):
    """`Qwen2_5_VLForConditionalGeneration.get_rope_index` (modeling_qwen2_5_vl_kv.py:1795-1990) for one sequence with
    images (no video, no padding): text tokens carry one running position on the three axes; the merged patches of an
    image get (t, h, w) grid offsets from the position where the image starts; text after it continues from max + 1.
    Returns (position_ids [3, 1, L], rope_delta = max + 1 - L), host-side like the reference."""
    ids = input_ids.reshape(-1).tolist()
    grids = [] if image_grid_thw is None else [[int(x) for x in g] for g in torch.as_tensor(image_grid_thw).tolist()]
    chunks, st = [], 0
    for t, h, w in grids:
        ed = ids.index(image_token_id, st)
        gh, gw = h // spatial_merge_size, w // spatial_merge_size
        st_idx = int(chunks[-1].max()) + 1 if chunks else 0
        text_len = ed - st
        chunks.append(torch.arange(text_len).view(1, -1).expand(3, -1) + st_idx)
        t_index = torch.arange(t).view(-1, 1).expand(-1, gh * gw).flatten() * 0   # images: second_per_grid_t = 0
        h_index = torch.arange(gh).view(1, -1, 1).expand(t, -1, gw).flatten()
        w_index = torch.arange(gw).view(1, 1, -1).expand(t, gh, -1).flatten()
        chunks.append(torch.stack([t_index, h_index, w_index]) + text_len + st_idx)
        st = ed + t * gh * gw
    if st < len(ids):
        st_idx = int(chunks[-1].max()) + 1 if chunks else 0
        chunks.append(torch.arange(len(ids) - st).view(1, -1).expand(3, -1) + st_idx)
    pos = torch.cat(chunks, dim=1).reshape(3, 1, -1)
    return pos, int(pos.max()) + 1 - len(ids)

