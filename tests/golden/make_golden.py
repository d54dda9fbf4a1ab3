#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED reference (KangJialiang/ViSpec d0240c87).

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports ``vispec.model.{cnets_ours,utils,kv_cache,configs,modeling_llama_kv}`` from
/root/reference with the four caller-side shims of SURVEY.md section 8(c) (nothing in the reference
is edited), loads the seeded synthetic weights of ``oracle/synth.py`` into the reference modules,
drives the reference's own cycle functions

    initialize_tree -> [tree_decoding -> evaluate_posterior -> update_inference_inputs] * n

and writes what they returned to ``tests/golden/*.pt``.  ``tests/test_oracle_golden.py`` then
checks ``oracle/vispec_oracle.py`` against these files; the GPU parity tests check the CUDA path
against the oracle and against these files.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

# shim 1: modeling_qwen2_5_vl_kv.py:34 wants a name transformers 5.x dropped
import transformers.cache_utils as _cu  # noqa: E402

if not hasattr(_cu, "SlidingWindowCache"):
    _cu.SlidingWindowCache = type("SlidingWindowCache", (), {})

from transformers import LlamaConfig  # noqa: E402
from vispec.model import cnets_ours, configs, kv_cache, modeling_llama_kv  # noqa: E402
from vispec.model import utils as ref_utils  # noqa: E402

from oracle import synth  # noqa: E402
from oracle.vispec_oracle import DraftCfg, TargetCfg  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def build_ref_target(tcfg: TargetCfg, weights):
    cfg = LlamaConfig(hidden_size=tcfg.hidden, intermediate_size=tcfg.inter, num_hidden_layers=tcfg.layers,
                      num_attention_heads=tcfg.heads, num_key_value_heads=tcfg.kv_heads, vocab_size=tcfg.vocab,
                      max_position_embeddings=tcfg.max_pos, rms_norm_eps=tcfg.eps)
    cfg.rope_scaling = None                     # shim 2
    cfg.rope_theta = tcfg.rope_theta
    cfg.architectures = ["LlamaForCausalLM"]    # shim 4
    cfg.pretraining_tp = 1
    m = modeling_llama_kv.LlamaForCausalLM(cfg).to(torch.bfloat16).eval()
    missing, unexpected = m.load_state_dict(weights, strict=False)
    assert not unexpected and all("rotary" in k for k in missing), (missing, unexpected)
    return m


def build_ref_draft(dcfg: DraftCfg, weights):
    ecfg = configs.EConfig(hidden_size=dcfg.hidden, intermediate_size=dcfg.inter, num_hidden_layers=1,
                           num_attention_heads=dcfg.heads, num_key_value_heads=dcfg.heads, vocab_size=dcfg.vocab,
                           max_position_embeddings=dcfg.max_pos, rms_norm_eps=dcfg.eps, pad_token_id=0)
    ecfg.rope_scaling = None                    # shim 2
    d = cnets_ours.Model(ecfg, bias=dcfg.bias, total_tokens=dcfg.total_tokens, depth=dcfg.depth,
                         top_k=dcfg.top_k, num_q=dcfg.num_q)
    missing, unexpected = d.load_state_dict(weights, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    d = d.to(torch.bfloat16).eval()
    d.init_tree()
    return d


class RefSpecStandIn(torch.nn.Module):
    """The 10-line stand-in for SpecModel (SURVEY.md 8c shim 4): SpecModel.__init__ needs a tokenizer
    and checkpoints on disk (spec_model_ours.py:55-57); its forward is restated from :205-245."""

    def __init__(self, base_model, spec_layer):
        super().__init__()
        self.base_model = base_model
        self.spec_layer = spec_layer

    def forward(self, input_ids=None, past_key_values=None, output_orig=False, position_ids=None,
                inputs_embeds=None, **kw):
        if inputs_embeds is not None:
            input_ids = None
        with torch.inference_mode():
            out = self.base_model(input_ids=input_ids, past_key_values=past_key_values, position_ids=position_ids,
                                  inputs_embeds=inputs_embeds, return_dict=True, output_hidden_states=True)
        return None, out.logits, out.hidden_states[-1]


def make_kv(tcfg: TargetCfg):
    """shim 3: initialize_past_key_values asserts CUDA (kv_cache.py:102); build the same structure by hand."""
    data = torch.zeros(2 * tcfg.layers, 1, tcfg.kv_heads, tcfg.max_pos, tcfg.head_dim, dtype=torch.bfloat16)
    lens = torch.zeros(2 * tcfg.layers, dtype=torch.long)
    pkv = [[kv_cache.KVCache(data[2 * i + j], lens[2 * i + j]) for j in (0, 1)] for i in range(tcfg.layers)]
    return pkv, [data], lens


@torch.no_grad()
def run_reference_cycles(model, tcfg, input_ids, inputs_embeds, image_mask, n_cycles):
    """spec_model_ours.py:455-547 at T=0, driven through the reference's own utils functions."""
    pkv, pkv_data, cur_len = make_kv(tcfg)
    model.spec_layer.reset_kv()
    ref_utils.reset_tree_mode(model)
    (draft_tokens, retrieve_indices, tree_mask, tree_position_ids, logits, hidden_state, sample_token) = \
        ref_utils.initialize_tree(input_ids, model, pkv, None, inputs_embeds, None, image_mask=image_mask)
    first = dict(draft_tokens=draft_tokens.clone(), retrieve_indices=retrieve_indices.clone(),
                 tree_mask=tree_mask.clone(), tree_position_ids=tree_position_ids.clone(),
                 sample_token=sample_token.clone(), prefill_last_logits=logits[0, -1].clone(),
                 draft_kv_len=model.spec_layer.stable_kv[0][0].shape[2],
                 draft_real_len=int(model.spec_layer.stable_kv[0][2]),
                 last_img_hidden=model.spec_layer.last_img_hidden.clone())
    new_token = 0
    padding = torch.zeros(1, 1, dtype=torch.long) - 1
    cycles = []
    for _ in range(n_cycles):
        model.base_model.model.tree_mask = tree_mask
        logits, hidden_state_new, _ = ref_utils.tree_decoding(model, draft_tokens, pkv, tree_position_ids,
                                                               input_ids, retrieve_indices)
        tree_logits_root = logits[0, 0].clone()
        draft_tokens = torch.cat((draft_tokens, padding), dim=1)
        candidates = draft_tokens[0, retrieve_indices]
        best_candidate, accept_length, sample_p = ref_utils.evaluate_posterior(logits, candidates, None)
        verify_argmax = torch.argmax(logits, dim=-1).clone()           # [n_leaves, depth]
        (input_ids, draft_tokens, retrieve_indices, tree_mask, tree_position_ids, new_token, _, sample_token) = \
            ref_utils.update_inference_inputs(input_ids, candidates, best_candidate, accept_length, retrieve_indices,
                                              None, new_token, pkv_data, cur_len, model, hidden_state_new, sample_p)
        cycles.append(dict(best_candidate=int(best_candidate), accept_length=int(accept_length),
                           candidates=candidates.clone(), verify_argmax=verify_argmax,
                           tree_logits_root=tree_logits_root, sample_token=sample_token.clone(),
                           draft_tokens=draft_tokens.clone(), retrieve_indices=retrieve_indices.clone(),
                           tree_mask=tree_mask.clone(), tree_position_ids=tree_position_ids.clone(),
                           kv_len=int(cur_len[0]), draft_kv_len=model.spec_layer.stable_kv[0][0].shape[2],
                           draft_real_len=int(model.spec_layer.stable_kv[0][2])))
    # checksum of the spliced slab (fp32 sum over the live prefix of every (layer, K/V))
    kv_sum = pkv_data[0][..., :int(cur_len[0]), :].float().sum(dim=(1, 2, 3, 4)).clone()
    return dict(first=first, cycles=cycles, output_ids=input_ids.clone(), new_token=int(new_token), kv_checksum=kv_sum)


@torch.no_grad()
def run_reference_greedy(base, tcfg, input_ids, inputs_embeds, n):
    """gen_baseline_answer_coco_caption.py:34-133 greedy loop over the same KVCache target."""
    pkv, _, _ = make_kv(tcfg)
    base.model.tree_mask = None
    out = base(inputs_embeds=inputs_embeds, past_key_values=pkv, use_cache=True) if inputs_embeds is not None \
        else base(input_ids=input_ids, past_key_values=pkv, use_cache=True)
    ids = input_ids.clone()
    for _ in range(n):
        tok = torch.argmax(out.logits[:, -1])[None, None]
        ids = torch.cat((ids, tok), dim=1)
        out = base(input_ids=tok, past_key_values=pkv, use_cache=True)
    return ids


CASES = {
    # name: (kind, seed, (n_pre, n_img, n_post), n_cycles, fc_noise)
    "random_img": ("random", 11, (5, 40, 12), 6, 0.0),
    "random_img_first": ("random", 12, (0, 24, 9), 4, 0.0),      # image run at position 0
    "random_text": ("random", 13, (17, 0, 0), 5, 0.0),           # no image: text-only branch (cnets_ours.py:977-981)
    "bigram_img": ("bigram", 21, (5, 40, 12), 6, 0.0),           # tau = depth+1
    "bigram_noisy_img": ("bigram", 22, (3, 32, 20), 8, 0.5),     # intermediate tau
}


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    tcfg, dcfg = synth.TINY_T, synth.TINY_D
    for name, (kind, seed, (n_pre, n_img, n_post), n_cycles, noise) in CASES.items():
        if kind == "random":
            tw = synth.make_target_weights(tcfg, seed)
            dw = synth.make_draft_weights(dcfg, seed + 1)
        else:
            tw, dw = synth.make_bigram_fixture(tcfg, dcfg, seed, fc_noise=noise)
        ids, emb, mask = synth.make_prompt(tcfg, tw, seed + 7, n_pre, n_img, n_post)
        base = build_ref_target(tcfg, tw)
        draft = build_ref_draft(dcfg, dw)
        model = RefSpecStandIn(base, draft)
        res = run_reference_cycles(model, tcfg, ids, emb if mask is not None else None, mask, n_cycles)
        n_new = res["output_ids"].shape[1] - ids.shape[1]
        greedy = run_reference_greedy(base, tcfg, ids, emb if mask is not None else None, n_new)
        res["greedy_ids"] = greedy
        res["spec_equals_greedy"] = bool(torch.equal(greedy, res["output_ids"]))
        res["meta"] = dict(kind=kind, seed=seed, n_pre=n_pre, n_img=n_img, n_post=n_post, n_cycles=n_cycles,
                           fc_noise=noise, torch=torch.__version__)
        torch.save(res, os.path.join(OUT, f"{name}.pt"))
        acc = [c["accept_length"] for c in res["cycles"]]
        print(f"{name}: accept={acc} new_token={res['new_token']} spec==greedy {res['spec_equals_greedy']} "
              f"retrieve={tuple(res['first']['retrieve_indices'].shape)}")


if __name__ == "__main__":
    main()
