"""ViSpec top-level model -- mirror of vispec/model/spec_model_ours.py `SpecModel` (EAGLE's `EaModel`).

Surface kept for vispec.evaluation (gen_spec_answer_*.py:103-116,146,153-154,173-178,224-230):
`SpecModel.from_pretrained(...)`, `.get_tokenizer()`, `.eval()`, `.training`, `.base_model`, `.spec_layer`,
`.specgenerate(input_ids, temperature=, log=, return_acceptance_len=, return_decode_time=, **processor_outputs)`.
`specgenerate` keeps the reference's return packing (spec_model_ours.py:555-582) and its stop rules
(:544-547); inside, each draft+verify cycle is ONE call into the C ABI (`vs_spec_cycle`) with no host sync,
and the host reads a 10-int record per cycle to apply the stop rules.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import time
from types import SimpleNamespace
from typing import Dict, Optional

import torch

from .. import _lib as L
from ..engine import DraftConfig, TargetConfig, TargetEngine, _stream, use_programs
from ..program import Program
from .cnets_ours import Model
from .configs import EConfig
from .kv_cache import initialize_past_key_values
from .utils import initialize_tree, prepare_logits_processor, reset_tree_mode


class _Weight:
    def __init__(self, weight):
        self.weight = weight


class _Inner:
    """`base_model.model`: carries tree_mask / tree_mode like modeling_llama_kv.LlamaModel (:917-922)."""

    def __init__(self, engine: TargetEngine):
        self.tree_mask = None
        self.tree_mode = None
        self.embed_tokens = engine.embed_tokens
        q = _Weight(engine.layers_t[-1]["w_qkv"])
        self.layers = [SimpleNamespace(self_attn=SimpleNamespace(q_proj=q)) for _ in range(engine.cfg.layers)]


class TargetLM:
    """`SpecModel.base_model` for a Llama-family target: callable like the reference's KV-cache LM
    (modeling_llama_kv.LlamaForCausalLM.forward :1118-1214) over the static slab."""

    def __init__(self, engine: TargetEngine, architectures=("LlamaForCausalLM",), image_token_index: Optional[int] = None):
        self.engine = engine
        self.config = SimpleNamespace(architectures=list(architectures), vocab_size=engine.cfg.vocab,
                                      hidden_size=engine.cfg.hidden, image_token_index=image_token_index,
                                      max_position_embeddings=engine.cfg.max_pos)
        self.lm_head = _Weight(engine.lm_head)
        self.model = _Inner(engine)
        self.dtype, self.device = engine.dtype, engine.device
        self.current_length_data = None        # set by initialize_past_key_values' caller (SpecModel)

    rope_deltas = 0
    # Prompt-pass logits: "last" returns [1, 1, V] (the row `initialize_tree` and the baseline loops consume through
    # `logits[:, -1]`); "all" computes the reference's [1, L, V] -- 185 GFLOP and 90 MB at 704 rows for one used row.
    prompt_logits = "last"

    def get_input_embeddings(self):
        return self.engine.embed_tokens

    def eval(self):
        return self

    def __call__(self, input_ids=None, attention_mask=None, past_key_values=None, position_ids=None,
                 inputs_embeds=None, use_cache=True, return_dict=True, output_hidden_states=False, **kwargs):
        eng = self.engine
        if past_key_values is None:
            raise ValueError("vispec_b200's target LM always runs over the static KV slab: pass past_key_values")
        if (input_ids is None) == (inputs_embeds is None):
            raise ValueError("You must specify exactly one of input_ids or inputs_embeds")
        length_t = past_key_values[0][0].current_length
        cur = int(length_t.item())
        eng.set_len(cur)
        q_len = inputs_embeds.shape[1] if inputs_embeds is not None else input_ids.shape[1]
        tree_mask = self.model.tree_mask
        if inputs_embeds is not None or q_len > L.VS_TOK_BLOCK:
            if tree_mask is not None:
                raise ValueError("tree-masked calls take at most 32 tokens")
            emb = inputs_embeds if inputs_embeds is not None else eng.embed_tokens(input_ids)
            if position_ids is not None and eng.cfg.mrope_section is not None and cur == 0:
                # Qwen2.5-VL: [3, 1, L] multimodal positions; rope_deltas = max + 1 - L (get_rope_index,
                # modeling_qwen2_5_vl_kv.py:1795-1990), cached on the model like the reference does (:2151-2164)
                p3 = position_ids.reshape(3, -1)
                self.rope_deltas = int(p3.max()) + 1 - p3.shape[1]
                eng.set_rope_delta(self.rope_deltas)
            logits_last, hidden = eng.prefill(emb, past=cur, position_ids=position_ids)
            if self.prompt_logits == "all":    # the reference's [1, L, V] (modeling_llama_kv.py:1196-1197)
                logits = torch.nn.functional.linear(hidden, eng.lm_head).float()
            else:                              # every caller on the path reads logits[:, -1] only (utils.py:284-290)
                logits = logits_last.reshape(1, 1, -1)
        else:
            ids = input_ids.to(eng.device, torch.int32).flatten().contiguous()
            if position_ids is None:       # modeling_qwen2_5_vl_kv.py:2166-2175: cache position + rope_deltas (0 for Llama)
                pos = torch.arange(cur, cur + q_len, dtype=torch.int32, device=eng.device) + eng.rope_delta
            else:                          # [3, 1, q] for Qwen2.5-VL: text tokens carry one id on all three axes
                pos = position_ids.to(eng.device, torch.int32).reshape(-1, q_len)[0].contiguous()
            bits = None
            if tree_mask is not None and q_len > 1:
                tm = tree_mask.to(eng.device).reshape(tree_mask.shape[-2], tree_mask.shape[-1]) != 0
                if tm.shape[0] != q_len:
                    raise ValueError("tree_mask size does not match the number of tree tokens")
                weights = (1 << torch.arange(q_len, device=eng.device, dtype=torch.int64))
                bits = (tm.long() * weights[None]).sum(dim=1).contiguous()
            eng.verify(ids, pos, bits, q_len, cur_len=cur)
            hidden = eng.hidden_out[:q_len].clone()[None]
            logits = eng.logits[:q_len].float()[None]
            eng.set_len(cur + q_len)
        base = length_t._base if length_t._base is not None else length_t
        base.fill_(cur + q_len)                     # every layer's KVCache shares this CPU tensor (kv_cache.py:130-132)
        return SimpleNamespace(logits=logits, hidden_states=(hidden,), past_key_values=past_key_values)


class SpecModel:
    def __init__(self, base_model: TargetLM, spec_layer: Model, tokenizer=None, eos_token_id: Optional[int] = None,
                 vision=None):
        self.base_model = base_model
        self.base_model_engine: TargetEngine = base_model.engine
        self.config = base_model.config
        self.spec_layer = spec_layer
        self.spec_layer.engine.bind_state(self.base_model_engine.state)
        self.spec_layer.engine.bind_tiled(self.base_model_engine.tiled)
        self.hidden_size = base_model.engine.cfg.hidden
        self.vocab_size = base_model.engine.cfg.vocab
        self.tokenizer = tokenizer if tokenizer is not None else SimpleNamespace(eos_token_id=eos_token_id)
        self.vision = vision          # callable(input_ids, inputs_embeds, **processor_outputs) -> (inputs_embeds, image_mask)
        self.training = False
        self.spec_layer.init_tree()
        dev = self.base_model_engine.device
        self._rec_host = [torch.zeros(2 + L.VS_MAX_ACCEPT, dtype=torch.int32).pin_memory() for _ in range(2)]
        self._rec_dev = [torch.zeros(2 + L.VS_MAX_ACCEPT, dtype=torch.int32, device=dev) for _ in range(2)]
        self._rec_evt = [torch.cuda.Event() for _ in range(2)]
        self._dec_evt = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        self.last_decode_ms = 0.0
        self._cycle_progs: Dict[tuple, Program] = {}
        self._accept_progs: Dict[float, Program] = {}

    # -- T > 0 (utils.py:453-493, 550-552) -----------------------------------------------------------------
    def sampling_setup(self, logits_processor, reseed: bool = False):
        """Device-side sampling state: temperature bits (the tree build sorts the leaf paths when set) and, once per
        generation, a Philox key drawn from torch's CPU generator (so `torch.manual_seed(i)` per prompt, as the
        evaluation scripts do, makes a generation reproducible)."""
        import struct
        st = self.base_model_engine.state
        if logits_processor is None:
            st[L.ST_TEMP_BITS] = 0
            return
        st[L.ST_TEMP_BITS] = struct.unpack("<i", struct.pack("<f", float(logits_processor.temperature)))[0]
        if reseed:
            st[L.ST_RNG_SEED] = int(torch.randint(0, 2 ** 31 - 1, (1,)))
            st[L.ST_RNG_CTR] = 0

    def accept_sampling_program(self, temperature: float, top_k: int = 0) -> Program:
        """The T > 0 acceptance + next-token sampling on the verify logits, for the stage-by-stage route."""
        p = self._accept_progs.get((temperature, top_k))
        if p is None:
            eng_t, eng_d = self.base_model_engine, self.spec_layer.engine
            p = Program(eng_t.device, top_k=top_k)
            p.add("accept_sampling", C.byref(eng_d.tb), L.ptr(eng_t.state), L.ptr(eng_t.logits), eng_t.cfg.vocab,
                  eng_t.cfg.vocab, C.c_float(temperature), None, L.ptr(eng_d.out_rec))
            self._accept_progs[(temperature, top_k)] = p.commit()
        return p

    def cycle_program(self, trace: bool = False, pf_ahead: Optional[int] = None, inflight: Optional[int] = None,
                      attn_fused: bool = True, temperature: float = 0.0, top_k: int = 0) -> Program:
        """The whole draft+verify cycle (spec_model_ours.py:484-547 loop body) as one persistent-kernel launch.
        attn_fused: single-phase attention on TMA-staged K/V (any context the scratch was sized for); False selects the
        two-phase A/B program, which needs a -DVS_TWO_PHASE_ATTN=1 build of the library."""
        eng_t, eng_d = self.base_model_engine, self.spec_layer.engine
        p = Program(eng_t.device, pf_ahead=pf_ahead, trace=trace, attn_fused=attn_fused, top_k=top_k)
        p.register_tiled(eng_t.tiled + eng_d.tiled)
        if inflight is not None:
            p.set_option("inflight", inflight)
        if temperature > 1e-5:
            p.add("spec_cycle_sampling", C.byref(eng_t.c), C.byref(eng_t.scratch.c), C.byref(eng_d.c),
                  C.byref(eng_d.scratch.c), C.byref(eng_d.tb), L.ptr(eng_t.state), L.ptr(eng_t.kv), eng_t.kv_max,
                  L.ptr(eng_t.hidden_out), L.ptr(eng_t.logits), C.c_float(temperature), L.ptr(self._rec_dev[0]))
        else:
            p.add("spec_cycle", C.byref(eng_t.c), C.byref(eng_t.scratch.c), C.byref(eng_d.c), C.byref(eng_d.scratch.c),
                  C.byref(eng_d.tb), L.ptr(eng_t.state), L.ptr(eng_t.kv), eng_t.kv_max, L.ptr(eng_t.hidden_out),
                  L.ptr(eng_t.logits), L.ptr(eng_t.amax), L.ptr(self._rec_dev[0]))
        return p.commit()

    # -- construction ----------------------------------------------------------------------------------
    @classmethod
    def from_state_dicts(cls, target_cfg: TargetConfig, target_weights: Dict[str, torch.Tensor],
                         draft_cfg: DraftConfig, draft_weights: Dict[str, torch.Tensor], device="cuda",
                         kv_max: Optional[int] = None, tokenizer=None, eos_token_id: Optional[int] = None,
                         image_token_index: Optional[int] = None, draft_cache_rows: int = 4096, vision=None):
        """Build from plain state dicts keyed like the reference modules (what tests and bench.py use: there are no
        checkpoints offline)."""
        eng = TargetEngine(target_cfg, target_weights, device=device, kv_max=kv_max)
        lm = TargetLM(eng, image_token_index=image_token_index)
        draft = Model(draft_cfg, draft_weights, eng.lm_head, device=device, cache_rows=draft_cache_rows,
                      max_positions=eng.kv_max + 64)
        return cls(lm, draft, tokenizer=tokenizer, eos_token_id=eos_token_id, vision=vision)

    @classmethod
    def from_pretrained(cls, Type="LLaMA", base_model_path=None, spec_model_path=None, total_token=30, depth=3,
                        top_k=8, threshold=1.0, num_q: int = 2, **kwargs):
        """spec_model_ours.py:110-203.  Loads a LLaVA(-Next)/Llama checkpoint with transformers, moves the language
        model into the B200 engine and keeps the HF vision tower + projector for the image features (host-side,
        out of the decode path).  Needs checkpoints on disk; not exercised offline."""
        from transformers import AutoConfig, AutoModelForCausalLM, AutoTokenizer

        arch = AutoConfig.from_pretrained(base_model_path).architectures[0]
        device = kwargs.pop("device", "cuda")
        kwargs.pop("device_map", None)
        kwargs.pop("low_cpu_mem_usage", None)
        kwargs.pop("torch_dtype", None)
        vision = None
        image_token_index = None
        if arch == "LlamaForCausalLM":
            hf = AutoModelForCausalLM.from_pretrained(base_model_path, torch_dtype=torch.bfloat16)
            sd, lcfg = hf.state_dict(), hf.config
        elif arch in ("LlavaNextForConditionalGeneration", "LlavaForConditionalGeneration"):
            from transformers import AutoModelForImageTextToText
            hf = AutoModelForImageTextToText.from_pretrained(base_model_path, torch_dtype=torch.bfloat16)
            lcfg = hf.config.text_config
            image_token_index = hf.config.image_token_index
            sd = {}
            for k, v in hf.state_dict().items():          # transformers 4.x and 5.x name the LM differently
                for pre in ("language_model.model.", "model.language_model."):
                    if k.startswith(pre):
                        sd["model." + k[len(pre):]] = v
                if k in ("language_model.lm_head.weight", "lm_head.weight"):
                    sd["lm_head.weight"] = v
            vision = _HFLlavaVision(hf, device)
        elif arch == "Qwen2_5_VLForConditionalGeneration":
            # spec_model_ours.py:137-140,380-449.  Needs a checkpoint on disk: not exercised offline (the Qwen target
            # itself is, at synthetic weights: tests/test_gpu_qwen.py).
            from transformers import AutoModelForImageTextToText
            hf = AutoModelForImageTextToText.from_pretrained(base_model_path, torch_dtype=torch.bfloat16)
            lcfg = getattr(hf.config, "text_config", hf.config)
            image_token_index = hf.config.image_token_id
            sd = {}
            for k, v in hf.state_dict().items():
                for pre in ("model.language_model.", "language_model.model."):
                    if k.startswith(pre):
                        sd["model." + k[len(pre):]] = v
                if k.startswith("model.layers.") or k in ("model.embed_tokens.weight", "model.norm.weight"):
                    sd[k] = v                         # transformers 4.49 layout (flat text model)
                if k in ("lm_head.weight", "language_model.lm_head.weight"):
                    sd["lm_head.weight"] = v
            if "lm_head.weight" not in sd:            # tied embeddings (3B)
                sd["lm_head.weight"] = sd["model.embed_tokens.weight"]
            vision = _HFQwenVision(hf.to(device))
            rs = getattr(lcfg, "rope_scaling", None) or getattr(lcfg, "rope_parameters", None) or {}
            qwen_extra = dict(mrope_section=tuple(rs.get("mrope_section", (16, 24, 24))), attn_impl="sdpa")
        else:
            raise NotImplementedError(f"Model type {arch} is not supported. Please use a supported model type.")
        rope_theta = getattr(lcfg, "rope_theta", None)
        if rope_theta is None:                        # transformers 5.x keeps it in rope_parameters
            rope_theta = (getattr(lcfg, "rope_parameters", None) or {}).get("rope_theta", 10000.0)
        tcfg = TargetConfig(hidden=lcfg.hidden_size, inter=lcfg.intermediate_size, layers=lcfg.num_hidden_layers,
                            heads=lcfg.num_attention_heads, kv_heads=lcfg.num_key_value_heads, vocab=lcfg.vocab_size,
                            max_pos=8192, eps=lcfg.rms_norm_eps, rope_theta=rope_theta,
                            **(qwen_extra if arch == "Qwen2_5_VLForConditionalGeneration" else {}))
        configpath = os.path.join(spec_model_path, "config.json")
        ecfg = EConfig.from_pretrained(configpath)
        dcfg = DraftConfig(hidden=ecfg.hidden_size, inter=ecfg.intermediate_size, heads=ecfg.num_attention_heads,
                           vocab=ecfg.vocab_size, max_pos=ecfg.max_position_embeddings, eps=ecfg.rms_norm_eps,
                           rope_theta=ecfg.rope_theta, total_tokens=total_token, depth=depth, top_k=top_k, num_q=num_q,
                           bias=ecfg.bias)
        p = os.path.join(spec_model_path, "pytorch_model.bin")
        if os.path.exists(p):
            dsd = torch.load(p, map_location="cpu")
        else:
            from safetensors.torch import load_file
            dsd = load_file(os.path.join(spec_model_path, "model.safetensors"))
        tok = AutoTokenizer.from_pretrained(base_model_path, use_fast=False)
        return cls.from_state_dicts(tcfg, sd, dcfg, dsd, device=device, tokenizer=tok,
                                    image_token_index=image_token_index, vision=vision)

    # -- nn.Module-ish surface the evaluation scripts touch ----------------------------------------------
    def enable_tensor_parallel_prefill(self, group=None):
        """Tensor-parallel target prefill over the ranks of `group` (BASELINE config 4; TargetEngine docstring).  All
        ranks then run specgenerate on the same prompt and produce the same tokens."""
        self.base_model_engine.enable_tensor_parallel_prefill(group)
        return self

    def get_tokenizer(self):
        return self.tokenizer

    def eval(self):
        self.training = False
        return self

    def __call__(self, *a, **k):
        return self.forward(*a, **k)

    def forward(self, input_ids=None, attention_mask=None, past_key_values=None, output_orig=False,
                position_ids=None, inputs_embeds=None, output_real_hidden=False, **kwargs):
        """spec_model_ours.py:205-245."""
        if inputs_embeds is not None:
            input_ids = None
        outputs = self.base_model(input_ids=input_ids, attention_mask=attention_mask, past_key_values=past_key_values,
                                  position_ids=position_ids, inputs_embeds=inputs_embeds, return_dict=True,
                                  output_hidden_states=True)
        hidden_states = outputs.hidden_states[-1]
        if output_real_hidden:
            return None, outputs.logits, hidden_states, outputs.hidden_states
        if output_orig:
            return None, outputs.logits, hidden_states
        return None, hidden_states

    # -- the generate loop ----------------------------------------------------------------------------------
    @torch.no_grad()
    def specgenerate(self, input_ids, temperature=0.0, top_p=0.0, top_k=0.0, max_new_tokens=512, max_length=2048,
                     log=False, is_llama3=False, inputs_embeds=None, return_acceptance_len=False,
                     return_decode_time=False, **kwargs):
        """spec_model_ours.py:248-582 at T=0."""
        if input_ids is None:
            raise ValueError("You must specify exactly one of input_ids or inputs_embeds")
        eng_t, eng_d = self.base_model_engine, self.spec_layer.engine
        dev = eng_t.device
        stop_token_id = self.tokenizer.convert_tokens_to_ids("<|eot_id|>") if is_llama3 else None
        max_length = max_length - self.spec_layer.total_tokens - 10
        logits_processor = prepare_logits_processor(temperature=temperature, top_p=top_p, top_k=top_k) \
            if temperature > 1e-5 else None
        input_ids = input_ids.clone().to(dev)
        self.spec_layer.reset_kv()
        if not hasattr(self, "past_key_values"):
            (self.past_key_values, self.past_key_values_data,
             self.current_length_data) = initialize_past_key_values(self.base_model)
        self.current_length_data.zero_()
        special_image_mask = kwargs.pop("image_mask", None)
        if self.vision is not None and kwargs.get("pixel_values") is not None:
            if inputs_embeds is not None:
                raise ValueError("You cannot specify both pixel_values and inputs_embeds at the same time, and must "
                                 "specify either one")
            front = self.vision(input_ids, **kwargs)
            inputs_embeds, special_image_mask = front[0], front[1]
            if len(front) > 2:                            # Qwen2.5-VL: multimodal prompt positions
                kwargs["position_ids"] = front[2]
        input_len = input_ids.shape[1]
        reset_tree_mode(self)
        self.sampling_setup(logits_processor, reseed=True)
        self._vs_seeded = True                           # initialize_tree reseeds only on the stage-by-stage route
        # Qwen2.5-VL target: [3, 1, L] multimodal prompt positions (what the reference's base model derives from
        # image_grid_thw, see `qwen_rope_index`); the decode positions are then offset by rope_deltas on the device
        self.base_model.rope_deltas = 0
        eng_t.set_rope_delta(0)
        try:
            initialize_tree(input_ids, self, self.past_key_values, logits_processor, inputs_embeds, None,
                            image_mask=special_image_mask, position_ids=kwargs.pop("position_ids", None))
        finally:
            self._vs_seeded = False
        new_token = 0
        acceptance_len = []
        if return_decode_time:
            torch.cuda.synchronize()
            start_time = time.time()
        eos = getattr(self.tokenizer, "eos_token_id", None)
        T = self.spec_layer.total_tokens + 1
        out_tokens = []
        cur_len = input_len
        d_rows = int(eng_t.state[L.ST_D_LEN].item())
        st = _stream()
        lib = eng_t.lib

        programs = use_programs()

        def launch(slot):
            if programs:
                # one cycle program serves every context length the scratch was sized for (the single-phase attention keeps
                # its scores in registers up to 4 granules per split and parks them beyond, csrc/attn_ring.cuh)
                fused = True
                temp = logits_processor.temperature if logits_processor is not None else 0.0
                tk = getattr(logits_processor, "top_k", 0) if logits_processor is not None else 0
                key = (fused, temp) if not tk else (fused, temp, tk)
                prog = self._cycle_progs.get(key)
                if prog is None:
                    prog = self._cycle_progs[key] = self.cycle_program(attn_fused=fused, temperature=temp, top_k=tk)
                prog.run(st)
                self._rec_host[slot].copy_(self._rec_dev[0], non_blocking=True)
                self._rec_evt[slot].record()
                return
            if logits_processor is not None:
                raise NotImplementedError("T > 0 runs on the persistent program path only (unset VS_MULTI_KERNEL)")
            L.check(lib.vs_spec_cycle(C.byref(eng_t.c), C.byref(eng_t.scratch.c), C.byref(eng_d.c),
                                      C.byref(eng_d.scratch.c), C.byref(eng_d.tb), L.ptr(eng_t.state), L.ptr(eng_t.kv),
                                      eng_t.kv_max, L.ptr(eng_t.hidden_out), L.ptr(eng_t.logits), L.ptr(eng_t.amax),
                                      L.ptr(self._rec_dev[slot]), st), "vs_spec_cycle")
            self._rec_host[slot].copy_(self._rec_dev[slot], non_blocking=True)
            self._rec_evt[slot].record()

        # the next cycle is launched before the previous record is read, so the host never idles the GPU;
        # a cycle launched past the stop point is simply ignored (its state is reset by the next call).
        margin = 2 * (T + L.VS_MAX_ACCEPT)
        idx = 0
        launched = 0

        def fits_now():
            # rows a cycle (and the one launched ahead of its record) may append: target slab, draft cache, and the
            # draft's rotary table (draft positions follow the uncompressed sequence index = cur_len)
            return (cur_len + margin + T < eng_t.kv_max and
                    d_rows + margin + 8 * (self.spec_layer.depth + 2) < eng_d.cache_rows and
                    cur_len + margin + self.spec_layer.depth + 2 < eng_d.rope_rows)

        self._dec_evt[0].record()
        if max_length > 0:
            if not fits_now():       # the reference raises in KVCache.cat (kv_cache.py:52-66); never launch past the slab
                raise ValueError(f"prompt of {input_len} tokens leaves no room for a draft tree in the KV slab "
                                 f"({eng_t.kv_max} rows) / draft cache ({eng_d.cache_rows} rows)")
            launch(0)
            launched = 1
        for idx in range(max_length):
            fits = fits_now()
            if launched == idx + 1 and idx + 1 < max_length and fits:
                launch((idx + 1) % 2)
                launched += 1
            slot = idx % 2
            self._rec_evt[slot].synchronize()
            rec = self._rec_host[slot].tolist()
            a = rec[0]
            out_tokens.extend(rec[2:3 + a])
            acceptance_len.append(a)
            new_token += a + 1
            cur_len += a + 1
            d_rows += a + 1
            if stop_token_id is not None and stop_token_id in out_tokens:
                break
            if eos is not None and eos in out_tokens:
                break
            if new_token > max_new_tokens:
                break
            if launched == idx + 1:          # ran out of cache room: cannot continue
                break
        self._dec_evt[1].record()
        self._dec_evt[1].synchronize()          # also drains a cycle launched past the stop point
        self.last_decode_ms = self._dec_evt[0].elapsed_time(self._dec_evt[1])
        input_ids = torch.cat([input_ids, torch.tensor([out_tokens], dtype=input_ids.dtype, device=dev)], dim=-1) \
            if out_tokens else input_ids
        self.current_length_data.fill_(input_ids.shape[1])
        outputs = (input_ids,)
        if log:
            outputs += (new_token, idx)
        if return_acceptance_len:
            outputs += (acceptance_len,)
        if return_decode_time:
            torch.cuda.synchronize()
            outputs += (time.time() - start_time,)
        return outputs[0] if len(outputs) == 1 else outputs

    # EAGLE-lineage aliases named by BASELINE.json's north_star
    ea_generate = specgenerate


EaModel = SpecModel


class _HFLlavaVision:
    """Host-side vision front-end (spec_model_ours.py:311-378): HF vision tower + projector + feature packing
    (spatial unpad + image_newline for LLaVA-Next).  Out of the decode path.  Only the vision tower, the projector, the
    newline row and the token embedding are kept on the device: the language model lives in the B200 engine."""

    def __init__(self, hf_model, device=None):
        self.m = hf_model
        self.cfg = hf_model.config
        inner = getattr(hf_model, "model", hf_model)
        lm = getattr(inner, "language_model", None)
        # free the HF language model's layers (a third copy of the LM weights); the embedding table stays
        self.embed = hf_model.get_input_embeddings()
        if lm is not None and hasattr(lm, "layers"):
            lm.layers = torch.nn.ModuleList()
        if hasattr(hf_model, "lm_head"):
            hf_model.lm_head = torch.nn.Identity()
        if device is not None:
            hf_model.to(device)

    def image_features(self, pixel_values, image_sizes):
        """[n_image_tokens, H] in the order of the image placeholders.  transformers 4.x returns per-image patch features
        and leaves the packing to the caller (`pack_image_features`, spec_model_ours.py:342-355); 5.x returns the packed
        rows in `pooler_output`.  Both are handled; keyword arguments only (the positional order changed)."""
        m, cfg = self.m, self.cfg
        kw = dict(vision_feature_layer=getattr(cfg, "vision_feature_layer", -2),
                  vision_feature_select_strategy=getattr(cfg, "vision_feature_select_strategy", "default"))
        if image_sizes is not None:
            out = m.get_image_features(pixel_values=pixel_values, image_sizes=image_sizes, **kw)
        else:                                         # LlavaForConditionalGeneration: no any-resolution packing
            out = m.get_image_features(pixel_values=pixel_values, **kw)
        feats = getattr(out, "pooler_output", out)
        if isinstance(feats, (list, tuple)):
            if feats and feats[0].dim() == 3 and image_sizes is not None:      # unpacked [patches, tokens, H] per image
                inner = getattr(m, "model", m)
                packer = m.pack_image_features if hasattr(m, "pack_image_features") else inner.pack_image_features
                feats, _ = packer(list(feats), image_sizes, vision_feature_select_strategy=kw["vision_feature_select_strategy"],
                                  image_newline=getattr(inner, "image_newline", getattr(m, "image_newline", None)))
            if isinstance(feats, (list, tuple)):
                feats = torch.cat([f.reshape(-1, f.shape[-1]) for f in feats], dim=0)
        return feats.reshape(-1, feats.shape[-1])

    @torch.no_grad()
    def __call__(self, input_ids, pixel_values=None, image_sizes=None, **kw):
        inputs_embeds = self.embed(input_ids)
        feats = self.image_features(pixel_values.to(inputs_embeds.device), image_sizes)
        mask = input_ids == self.cfg.image_token_index
        if int(mask.sum()) != feats.shape[0]:
            raise ValueError(f"Image features and image tokens do not match: tokens: {int(mask.sum())}, "
                             f"features {feats.shape[0]}")
        inputs_embeds = inputs_embeds.masked_scatter(mask.unsqueeze(-1).expand_as(inputs_embeds),
                                                     feats.to(inputs_embeds.device, inputs_embeds.dtype))
        return inputs_embeds, mask


class _HFQwenVision:
    """Host-side front-end of a Qwen2.5-VL checkpoint (spec_model_ours.py:380-417 + get_rope_index): token embeddings,
    vision tower over `pixel_values` / `image_grid_thw`, scatter into the image placeholders, multimodal prompt
    positions.  Out of the decode path; requires a real checkpoint (images only, no video)."""

    def __init__(self, hf_model):
        self.m = hf_model

    @torch.no_grad()
    def __call__(self, input_ids, pixel_values=None, image_grid_thw=None, **kw):
        from .utils import qwen_rope_index
        m = self.m
        inputs_embeds = m.get_input_embeddings()(input_ids)
        visual = m.visual if hasattr(m, "visual") else m.model.visual
        image_embeds = visual(pixel_values.type(visual.dtype), grid_thw=image_grid_thw)
        image_embeds = getattr(image_embeds, "pooler_output", image_embeds)
        image_token_id = m.config.image_token_id
        mask = input_ids == image_token_id
        if int(mask.sum()) != image_embeds.shape[0]:
            raise ValueError(f"Image features and image tokens do not match: tokens: {int(mask.sum())}, "
                             f"features {image_embeds.shape[0]}")
        inputs_embeds = inputs_embeds.masked_scatter(mask.unsqueeze(-1).expand_as(inputs_embeds),
                                                     image_embeds.to(inputs_embeds.device, inputs_embeds.dtype))
        merge = getattr(m.config.vision_config, "spatial_merge_size", 2)
        pos3, _ = qwen_rope_index(input_ids.cpu(), image_grid_thw=image_grid_thw.cpu(), image_token_id=image_token_id,
                                  spatial_merge_size=merge)
        return inputs_embeds, mask, pos3.to(input_ids.device)

