"""Drive the UNMODIFIED reference (KangJialiang/ViSpec d0240c87) from its vendored copy under baseline/_ref/.

This is the reference arm of bench.py (`--impl reference` on the host cores, `--impl reference-gpu` and the
`gpu_eager_baseline` leg on the B200) and of the drop-in test.  Nothing in here is product code and nothing under
vispec_b200/ imports it.  baseline/_ref/ is a plain copy made by tools/vendor_reference.py (the reference has no
build system), so the GPU box -- which has no /root/reference -- runs the very same files.

What runs is the reference's own hot path: `vispec.model.modeling_llama_kv.LlamaForCausalLM` (target with the static
KV slab of kv_cache.py), `vispec.model.cnets_ours.Model` (draft head), and the cycle functions of
`vispec.model.utils` (`initialize_tree`, `tree_decoding`, `evaluate_posterior`, `update_inference_inputs`) in the
order `SpecModel.specgenerate` calls them (spec_model_ours.py:455-547), timed over the region the reference itself times
(`return_decode_time`, spec_model_ours.py:480-482,575-578).  `SpecModel.__init__` needs a tokenizer and checkpoints on
disk (spec_model_ours.py:55-57), which do not exist offline, so a 10-line stand-in with the same `forward` return
packing (spec_model_ours.py:205-245) holds the two modules; the four caller-side shims are those of SURVEY.md 8c.
"""
from __future__ import annotations

import os
import sys
import time
from typing import Dict, Optional

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref")
_mods = None


def available() -> Optional[str]:
    """None when the vendored reference is importable, else the reason (one line)."""
    if not os.path.exists(os.path.join(REF_ROOT, "vispec", "model", "cnets_ours.py")):
        return "baseline/_ref is empty: run tools/vendor_reference.py where /root/reference exists"
    try:
        load()
    except Exception as e:  # pragma: no cover - depends on the environment
        return f"reference import failed: {type(e).__name__}: {e}"
    return None


def load():
    """Import the reference modules (shim 1: a name transformers 5.x dropped, modeling_qwen2_5_vl_kv.py:34)."""
    global _mods
    if _mods is not None:
        return _mods
    import transformers.cache_utils as cu
    if not hasattr(cu, "SlidingWindowCache"):
        cu.SlidingWindowCache = type("SlidingWindowCache", (), {})
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    stale = [k for k in sys.modules if k == "vispec" or k.startswith("vispec.")]
    for k in stale:                                   # a test may have aliased vispec.model to the mirror
        del sys.modules[k]
    from vispec.model import cnets_ours, configs, kv_cache, modeling_llama_kv
    from vispec.model import utils as ref_utils

    class M:
        pass
    m = M()
    m.cnets_ours, m.configs, m.kv_cache, m.modeling_llama_kv, m.utils = cnets_ours, configs, kv_cache, modeling_llama_kv, ref_utils
    _mods = m
    return m


class RefSpecStandIn(torch.nn.Module):
    """Stand-in for SpecModel (forward restated from spec_model_ours.py:205-245; see the module docstring)."""

    def __init__(self, base_model, spec_layer):
        super().__init__()
        self.base_model = base_model
        self.spec_layer = spec_layer

    def forward(self, input_ids=None, past_key_values=None, output_orig=False, position_ids=None, inputs_embeds=None, **kw):
        if inputs_embeds is not None:
            input_ids = None
        with torch.inference_mode():
            out = self.base_model(input_ids=input_ids, past_key_values=past_key_values, position_ids=position_ids,
                                  inputs_embeds=inputs_embeds, return_dict=True, output_hidden_states=True)
        return None, out.logits, out.hidden_states[-1]


def _construct(fn, device):
    """Build a reference module directly in bf16 on `device` without running the (discarded) random init.  The rotary
    tables are computed in fp32 and cast once either way (modeling_llama_kv.py:147-187), so this equals fp32
    construction followed by `.to(bfloat16)` (what tests/golden/make_golden.py does)."""
    try:
        from transformers.initialization import no_init_weights
    except ImportError:  # pragma: no cover
        import contextlib
        no_init_weights = contextlib.nullcontext
    old = torch.get_default_dtype()
    saved = {n: getattr(torch.nn.init, n) for n in ("kaiming_uniform_", "uniform_", "normal_", "trunc_normal_")}
    try:
        torch.set_default_dtype(torch.bfloat16)
        for n in saved:                                # nn.Linear / nn.Embedding reset_parameters: weights are overwritten
            setattr(torch.nn.init, n, lambda t, *a, **k: t)
        with torch.device(device), no_init_weights():
            return fn()
    finally:
        torch.set_default_dtype(old)
        for n, f in saved.items():
            setattr(torch.nn.init, n, f)


def build_target(tcfg, weights: Dict[str, torch.Tensor], device, assign: bool = False):
    """`assign`: the module's parameters alias the tensors of `weights` (no second copy of a 7B model)."""
    from transformers import LlamaConfig
    m = load()
    cfg = LlamaConfig(hidden_size=tcfg.hidden, intermediate_size=tcfg.inter, num_hidden_layers=tcfg.layers,
                      num_attention_heads=tcfg.heads, num_key_value_heads=tcfg.kv_heads, vocab_size=tcfg.vocab,
                      max_position_embeddings=tcfg.max_pos, rms_norm_eps=tcfg.eps)
    cfg.rope_scaling = None                     # shim 2 (cnets_ours.py:299 / modeling_llama_kv.py:500-503)
    cfg.rope_theta = tcfg.rope_theta
    cfg.architectures = ["LlamaForCausalLM"]    # shim 4 (utils.py:398)
    cfg.pretraining_tp = 1
    model = _construct(lambda: m.modeling_llama_kv.LlamaForCausalLM(cfg), device).eval()
    missing, unexpected = model.load_state_dict(weights, strict=False, assign=assign)
    assert not unexpected and all("rotary" in k for k in missing), (missing, unexpected)
    return model


def build_draft(dcfg, weights: Dict[str, torch.Tensor], device, assign: bool = False):
    m = load()
    ecfg = m.configs.EConfig(hidden_size=dcfg.hidden, intermediate_size=dcfg.inter, num_hidden_layers=1,
                             num_attention_heads=dcfg.heads, num_key_value_heads=dcfg.heads, vocab_size=dcfg.vocab,
                             max_position_embeddings=dcfg.max_pos, rms_norm_eps=dcfg.eps, pad_token_id=0)
    ecfg.rope_scaling = None                    # shim 2
    d = _construct(lambda: m.cnets_ours.Model(ecfg, bias=dcfg.bias, total_tokens=dcfg.total_tokens, depth=dcfg.depth,
                                              top_k=dcfg.top_k, num_q=dcfg.num_q), device)
    missing, unexpected = d.load_state_dict(weights, strict=False, assign=assign)
    assert not missing and not unexpected, (missing, unexpected)
    d = d.eval()
    d.init_tree()
    return d


def make_kv(model, tcfg, device):
    """KV slab + length tensors.  On CUDA this is the reference's own allocator; on CPU it asserts (kv_cache.py:102), so
    the same structure is built by hand (shim 3)."""
    m = load()
    if torch.device(device).type == "cuda":
        return m.kv_cache.initialize_past_key_values(model.base_model)
    data = torch.zeros(2 * tcfg.layers, 1, tcfg.kv_heads, tcfg.max_pos, 128, dtype=torch.bfloat16)
    lens = torch.zeros(2 * tcfg.layers, dtype=torch.long)
    pkv = [[m.kv_cache.KVCache(data[2 * i + j], lens[2 * i + j]) for j in (0, 1)] for i in range(tcfg.layers)]
    return pkv, [data], lens


def _sync(device):
    if torch.device(device).type == "cuda":
        torch.cuda.synchronize(device)


@torch.no_grad()
def generate(model, kv, input_ids, inputs_embeds, image_mask, max_new_tokens: int, max_cycles: int = 10 ** 9,
             synthetic_prefill: bool = False, count_launches: bool = False):
    """spec_model_ours.py:248-582 at T=0 through the reference's own functions.  `kv` = make_kv(...) (reused across calls
    like SpecModel caches it, :286-307).  Returns a dict: output_ids, new_token, accept (list), decode_s (the
    `return_decode_time` region), prefill_s, cycles, and launches_per_cycle when `count_launches` (CUDA only).

    `synthetic_prefill`: skip the target's prompt pass (minutes on CPU at 7B) -- the KV slab rows [0, L) are filled with
    random values, the first token is random and the draft is prefilled from random hidden states; every cycle after
    that does exactly the work of a real one ("decode only")."""
    m = load()
    ru = m.utils
    pkv, pkv_data, cur_len = kv
    device = input_ids.device
    cur_len.zero_()
    model.spec_layer.reset_kv()
    ru.reset_tree_mode(model)
    _sync(device)
    t0 = time.perf_counter()
    if synthetic_prefill:
        L = input_ids.shape[1]
        g = torch.Generator(device="cpu").manual_seed(1234)
        for d in pkv_data:
            d[..., :L, :] = torch.randn(d.shape[0], 1, d.shape[2], L, d.shape[4], generator=g).to(d.dtype).to(d.device)
        cur_len.fill_(L)
        H = model.base_model.config.hidden_size
        hidden = torch.randn(1, L, H, generator=g).to(torch.bfloat16).to(device)
        token = torch.randint(1, model.base_model.config.vocab_size - 64, (1, 1), generator=g).to(device)
        ids1 = torch.cat((input_ids, token), dim=1)
        tree = model.spec_layer.topK_genrate(hidden, ids1, model.base_model.lm_head, None, inputs_embeds=inputs_embeds,
                                             image_mask=image_mask)
        draft_tokens, retrieve_indices, tree_mask, tree_position_ids = tree
    else:
        (draft_tokens, retrieve_indices, tree_mask, tree_position_ids, logits, hidden_state, sample_token) = \
            ru.initialize_tree(input_ids, model, pkv, None, inputs_embeds, None, image_mask=image_mask)
    _sync(device)
    t1 = time.perf_counter()                                   # spec_model_ours.py:480-482
    new_token = 0
    padding = (torch.zeros(1, 1, dtype=torch.long) - 1).to(device)
    acc = []
    prof = None
    launches = None
    n_cycles = 0
    for idx in range(max_cycles):
        if count_launches and idx == 2:
            from torch.profiler import ProfilerActivity, profile
            prof = profile(activities=[ProfilerActivity.CUDA])
            prof.__enter__()
        model.base_model.model.tree_mask = tree_mask
        draft_tokens = draft_tokens.to(device)
        logits, hidden_state_new, _ = ru.tree_decoding(model, draft_tokens, pkv, tree_position_ids, input_ids,
                                                       retrieve_indices)
        draft_tokens = torch.cat((draft_tokens, padding), dim=1)
        candidates = draft_tokens[0, retrieve_indices]
        best_candidate, accept_length, sample_p = ru.evaluate_posterior(logits, candidates, None)
        acc.append(int(accept_length))
        (input_ids, draft_tokens, retrieve_indices, tree_mask, tree_position_ids, new_token, _, sample_token) = \
            ru.update_inference_inputs(input_ids, candidates, best_candidate, accept_length, retrieve_indices, None,
                                       new_token, pkv_data, cur_len, model, hidden_state_new, sample_p)
        n_cycles += 1
        if prof is not None and idx == 3:
            _sync(device)
            prof.__exit__(None, None, None)
            n_k = sum(1 for e in prof.events() if getattr(e, "device_type", None) is not None
                      and "cuda" in str(e.device_type).lower())
            launches = n_k / 2.0
            prof = None
        if new_token > max_new_tokens:
            break
    _sync(device)
    t2 = time.perf_counter()
    if prof is not None:
        prof.__exit__(None, None, None)
    return dict(output_ids=input_ids, new_token=int(new_token), accept=acc, decode_s=t2 - t1, prefill_s=t1 - t0,
                cycles=n_cycles, launches_per_cycle=launches)
