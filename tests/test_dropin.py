"""Drop-in proof: the reference's OWN evaluation code drives vispec_b200 through the names it imports.

  * CPU: the host-side LLaVA-Next vision front-end (`_HFLlavaVision`, spec_model_ours.py:311-378) against the merge HF's
    own model performs, on a tiny random LLaVA-Next checkpoint (the path needs no GPU: it is outside the decode path).
  * GPU: `vispec.evaluation.gen_spec_answer_coco_caption.get_model_answers` -- the unmodified reference file, vendored under
    baseline/_ref by tools/vendor_reference.py -- runs with `vispec.model` aliased to `vispec_b200.model`:
    `SpecModel.from_pretrained` on tiny checkpoints saved to disk (LLaVA-Next target incl. its key mapping, draft head
    `config.json` + `pytorch_model.bin`), `model.get_tokenizer()`, `.eval()`, `specgenerate(**processor_outputs,
    temperature=, log=True[, return_acceptance_len=True])`, and the JSONL record schema of :262-299.
Only third-party modules that are not installed offline are stubbed (shortuuid, datasets, the HF hub processor)."""
import copy
import importlib.util
import json
import os
import sys
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def tiny_llava_next(vocab=1000, image_token=999):
    from transformers import CLIPVisionConfig, LlamaConfig, LlavaNextConfig, LlavaNextForConditionalGeneration
    vc = CLIPVisionConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=2, image_size=28,
                          patch_size=14)
    tc = LlamaConfig(hidden_size=256, intermediate_size=640, num_hidden_layers=2, num_attention_heads=2,
                     num_key_value_heads=2, vocab_size=vocab, max_position_embeddings=512, rms_norm_eps=1e-5)
    cfg = LlavaNextConfig(vision_config=vc, text_config=tc, image_token_index=image_token,
                          image_grid_pinpoints=[[28, 28], [56, 28], [28, 56]], vision_feature_layer=-2)
    torch.manual_seed(0)
    m = LlavaNextForConditionalGeneration(cfg).eval()
    with torch.no_grad():                       # HF's sigma = 0.02 init makes every bf16 log-prob tie
        m.get_input_embeddings().weight.normal_(0, 1.0)
        m.lm_head.weight.normal_(0, 4.0 / 16.0)
    return m


def tiny_prompt(m, n_pre=4, n_post=9, seed=0):
    """input_ids with exactly as many image placeholders as the model packs features for one 50 x 28 image."""
    g = torch.Generator().manual_seed(seed)
    pv = torch.randn(1, 3, 3, 28, 28, generator=g)             # base view + 2 patches of the 56 x 28 grid
    sizes = torch.tensor([[50, 28]])
    with torch.no_grad():
        out = m.get_image_features(pixel_values=pv, image_sizes=sizes, vision_feature_layer=m.config.vision_feature_layer,
                                   vision_feature_select_strategy=m.config.vision_feature_select_strategy)
    feats = getattr(out, "pooler_output", out)
    n_img = sum(f.shape[0] for f in feats) if isinstance(feats, (list, tuple)) else feats.shape[0]
    ids = torch.randint(1, 900, (1, n_pre + n_img + n_post), generator=g)
    ids[0, n_pre:n_pre + n_img] = m.config.image_token_index
    return ids, pv, sizes, n_img


def test_llava_next_vision_front_end_matches_hf_merge():
    from vispec_b200.model.spec_model_ours import _HFLlavaVision
    m = tiny_llava_next()
    ids, pv, sizes, n_img = tiny_prompt(m)
    seen = {}
    lm = m.model.language_model
    h = lm.register_forward_pre_hook(lambda mod, args, kwargs: seen.update(emb=kwargs["inputs_embeds"].detach().clone()),
                                     with_kwargs=True)
    with torch.no_grad():
        m(input_ids=ids, pixel_values=pv, image_sizes=sizes)
    h.remove()
    front = _HFLlavaVision(copy.deepcopy(m))
    emb, mask = front(ids, pixel_values=pv, image_sizes=sizes)
    assert int(mask.sum()) == n_img and torch.equal(mask, ids == m.config.image_token_index)
    torch.testing.assert_close(emb, seen["emb"], rtol=0, atol=0)
    assert len(getattr(front.m.model.language_model, "layers")) == 0        # the HF language model was dropped
    with pytest.raises(ValueError, match="do not match"):                    # spec_model_ours.py:363-370
        front(ids[:, :-12], pixel_values=pv, image_sizes=sizes) if n_img > 12 else front(ids[:, 5:], pixel_values=pv, image_sizes=sizes)


def _save_checkpoints(tmp, m):
    from tokenizers import Tokenizer, models, pre_tokenizers
    from transformers import PreTrainedTokenizerFast
    base, spec = os.path.join(tmp, "base"), os.path.join(tmp, "spec")
    os.makedirs(spec)
    m.to(torch.bfloat16).save_pretrained(base)
    vocab = {f"t{i}": i for i in range(1000)}
    tok = Tokenizer(models.WordLevel(vocab=vocab, unk_token="t0"))
    tok.pre_tokenizer = pre_tokenizers.Whitespace()
    PreTrainedTokenizerFast(tokenizer_object=tok, unk_token="t0", bos_token="t1", eos_token="t2").save_pretrained(base)
    from oracle import synth
    dc = synth.TINY_D
    dw = synth.make_draft_weights(dc, 99)
    torch.save(dw, os.path.join(spec, "pytorch_model.bin"))
    with open(os.path.join(spec, "config.json"), "w") as f:      # the keys of vispec/train/llava_1.6_7B_config.json
        json.dump({"architectures": ["LlavaNextForConditionalGeneration"], "hidden_size": dc.hidden,
                   "intermediate_size": dc.inter, "max_position_embeddings": dc.max_pos, "num_attention_heads": dc.heads,
                   "num_hidden_layers": 1, "num_key_value_heads": dc.heads, "rms_norm_eps": dc.eps, "vocab_size": dc.vocab,
                   "hidden_act": "silu", "model_type": "llava_next"}, f)
    return base, spec


@pytest.mark.gpu
def test_reference_evaluation_script_runs_on_the_mirror(tmp_path, monkeypatch):
    script = os.path.join(REF, "vispec", "evaluation", "gen_spec_answer_coco_caption.py")
    if not os.path.exists(script):
        pytest.skip("baseline/_ref not vendored (tools/vendor_reference.py needs /root/reference)")
    m = tiny_llava_next()
    ids, pv, sizes, n_img = tiny_prompt(m)
    base, spec = _save_checkpoints(str(tmp_path), m)
    # ---- module graph the script imports: vispec.model.* -> the mirror; third-party packages missing offline -> stubs
    import vispec_b200.model as mirror
    import vispec_b200.model.kv_cache
    import vispec_b200.model.spec_model_ours
    import vispec_b200.model.utils
    pkg = types.ModuleType("vispec"); pkg.__path__ = []
    ev = types.ModuleType("vispec.evaluation"); ev.__path__ = []
    for k in [k for k in sys.modules if k == "vispec" or k.startswith("vispec.")]:
        monkeypatch.delitem(sys.modules, k)
    monkeypatch.setitem(sys.modules, "vispec", pkg)
    monkeypatch.setitem(sys.modules, "vispec.evaluation", ev)
    monkeypatch.setitem(sys.modules, "vispec.model", mirror)
    for sub in ("kv_cache", "spec_model_ours", "utils"):
        monkeypatch.setitem(sys.modules, "vispec.model." + sub, getattr(mirror, sub))
    shortuuid = types.ModuleType("shortuuid"); shortuuid.uuid = lambda: "stub-uuid"
    datasets = types.ModuleType("datasets"); datasets.Dataset = object; datasets.load_dataset = lambda *a, **k: None
    monkeypatch.setitem(sys.modules, "shortuuid", shortuuid)
    monkeypatch.setitem(sys.modules, "datasets", datasets)
    prompt = types.ModuleType("vispec.evaluation.coco_caption_prompt")

    def build_prompt(d, args):      # what processor(images=, text=, return_tensors="pt").to("cuda:0") returns
        return {"input_ids": ids.to("cuda:0"), "attention_mask": torch.ones_like(ids).to("cuda:0"),
                "pixel_values": pv.to("cuda:0", torch.bfloat16), "image_sizes": sizes.to("cuda:0")}

    prompt.build_prompt = build_prompt
    monkeypatch.setitem(sys.modules, "vispec.evaluation.coco_caption_prompt", prompt)
    spec_mod = importlib.util.spec_from_file_location("vispec.evaluation.gen_spec_answer_coco_caption", script)
    gen = importlib.util.module_from_spec(spec_mod)
    monkeypatch.setitem(sys.modules, spec_mod.name, gen)
    spec_mod.loader.exec_module(gen)                 # the reference file itself, unmodified
    # ---- run the reference's get_model_answers body on two "questions"
    args = types.SimpleNamespace(use_ours=True, use_medusa=False, total_token=30, depth=3, top_k=8, num_q=2, model=base)
    answer_file = os.path.join(str(tmp_path), "out", "answers.jsonl")
    data = [{"imgid": 7, "image": None}, {"imgid": 3, "image": None}]
    gen.get_model_answers(base, spec, "tiny-llava-next", data, answer_file, 64, 1, 1, None, 0.0, args)
    recs = [json.loads(line) for line in open(answer_file)]
    assert [r["question_id"] for r in recs] == [7, 3]
    for r in recs:
        assert set(r) == {"question_id", "answer_id", "model_id", "choices", "tstamp"} and r["model_id"] == "tiny-llava-next"
        (c,) = r["choices"]
        assert set(c) == {"index", "turns", "idxs", "new_tokens", "wall_time", "acceptance_length"}
        assert len(c["turns"]) == len(c["idxs"]) == len(c["new_tokens"]) == len(c["wall_time"]) == 1
        assert c["new_tokens"][0] == sum(c["acceptance_length"]) + len(c["acceptance_length"]) > 0
        assert c["idxs"][0] == len(c["acceptance_length"]) - 1 and isinstance(c["turns"][0], str)
    assert recs[0]["choices"][0]["turns"] == recs[1]["choices"][0]["turns"]       # same prompt, T = 0: same text
    gen.reorg_answer_file(answer_file)
    assert [json.loads(line)["question_id"] for line in open(answer_file)] == [3, 7]
    # the result reducer of the reference's speed.py consumes these records (vispec_b200.report mirrors it)
    from vispec_b200 import report
    assert report is not None
