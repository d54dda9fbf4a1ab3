"""CPU oracle for the ViSpec draft-and-verify decode path.  TEST INFRASTRUCTURE ONLY.

This module is a plain-PyTorch restatement of the reference algorithm
(KangJialiang/ViSpec, commit d0240c87).  It exists so that the CUDA product path in
``vispec_b200`` can be checked against something that runs anywhere torch runs.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu-baseline / reference
legs may import it; nothing under ``vispec_b200/`` does.

Parity status: PINNED against the reference itself.  ``tests/golden/make_golden.py`` imports
the unmodified reference modules from /root/reference (with the four caller-side shims of
SURVEY.md section 8c), runs them on seeded synthetic weights and commits the outputs under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks this restatement against those
files bit-for-bit (integers) / exactly (bf16 on CPU, same ATen ops in the same order).  The
reference ships no tests or golden vectors of its own.

Every function cites the reference file:line it follows (paths relative to the reference root).
Weights are plain ``dict[str, Tensor]`` keyed exactly like the reference ``state_dict``s.
Batch size is 1 throughout, as in the reference (kv_cache.py:88, cnets_ours.py:901-902).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

F32_MIN = torch.finfo(torch.float32).min


# --------------------------------------------------------------------------------------
# configs
# --------------------------------------------------------------------------------------
@dataclass
class TargetCfg:
    """Shape of the target LM (vispec/model/modeling_llama_kv.py LlamaConfig fields)."""
    hidden: int = 4096
    inter: int = 11008
    layers: int = 32
    heads: int = 32
    kv_heads: int = 32
    vocab: int = 32064
    max_pos: int = 8192          # forced by modeling_llava_next_kv.py:10
    eps: float = 1e-5
    rope_theta: float = 10000.0
    # Qwen2.5-VL target (modeling_qwen2_5_vl_kv.py): q/k/v biases (:830-838), multimodal RoPE sections over head_dim/2
    # (:740-782), SDPA attention with a model-dtype mask as the default attention class (:1065-1170)
    qkv_bias: bool = False
    mrope_section: Optional[tuple] = None
    attn_impl: str = "eager"

    @property
    def head_dim(self) -> int:
        return self.hidden // self.heads


@dataclass
class DraftCfg:
    """Shape of the draft model (vispec/train/llava_1.6_7B_config.json + exp.sh:11-14)."""
    hidden: int = 4096
    inter: int = 11008
    heads: int = 32
    vocab: int = 32064
    max_pos: int = 4096
    eps: float = 1e-5
    rope_theta: float = 10000.0   # cnets_ours.py:295-297 default base
    total_tokens: int = 30        # Model stores total_tokens-1 (cnets_ours.py:733)
    depth: int = 3
    top_k: int = 8
    num_q: int = 2
    bias: bool = True

    @property
    def head_dim(self) -> int:
        return self.hidden // self.heads


# --------------------------------------------------------------------------------------
# shared building blocks
# --------------------------------------------------------------------------------------
def rope_tables(dim: int, max_pos: int, base: float, dtype: torch.dtype, device="cpu"):
    """cos/sin tables [max_pos, dim]: fp32 math then ONE cast to the model dtype.

    cnets_ours.py:129-155 / modeling_llama_kv.py:153-187 (buffers are cast when the module is
    ``.to(dtype)``-ed, spec_model_ours.py:98)."""
    inv_freq = 1.0 / (base ** (torch.arange(0, dim, 2).float() / dim))
    t = torch.arange(max_pos, dtype=inv_freq.dtype)
    freqs = torch.einsum("i,j->ij", t, inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype).to(device), emb.sin().to(dtype).to(device)


def _rotate_half(x):
    """cnets_ours.py:104-108."""
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def apply_rope(q, k, cos, sin, position_ids):
    """cnets_ours.py:111-119 / modeling_llama_kv.py:338-358.  Three bf16 roundings each."""
    c = cos[position_ids].unsqueeze(1)
    s = sin[position_ids].unsqueeze(1)
    return (q * c) + (_rotate_half(q) * s), (k * c) + (_rotate_half(k) * s)


def apply_mrope(q, k, cos, sin, position_ids3, mrope_section):
    """modeling_qwen2_5_vl_kv.py:740-782 (`apply_multimodal_rotary_pos_emb`): position_ids3 is [3, q] (temporal, height,
    width); channel chunk i of cos / sin (sections doubled over the two halves) comes from axis i % 3.  The reference
    computes cos / sin on the fly in fp32 and casts once (:690-718) -- the same values as the table rows."""
    c3, s3 = cos[position_ids3], sin[position_ids3]                      # [3, q, D]
    sec = list(mrope_section) * 2
    c = torch.cat([m[i % 3] for i, m in enumerate(c3.split(sec, dim=-1))], dim=-1)[None, None]
    s = torch.cat([m[i % 3] for i, m in enumerate(s3.split(sec, dim=-1))], dim=-1)[None, None]
    return (q * c) + (_rotate_half(q) * s), (k * c) + (_rotate_half(k) * s)


def rms_norm(x, weight, eps):
    """cnets_ours.py:522-527 / modeling_llama_kv.py:119-133: fp32 normalise, cast, then * weight."""
    dt = x.dtype
    xf = x.to(torch.float32)
    var = xf.pow(2).mean(-1, keepdim=True)
    xf = xf * torch.rsqrt(var + eps)
    return weight * xf.to(dt)


def causal_mask(q_len: int, past: int, device) -> torch.Tensor:
    """fp32 additive causal mask [1,1,q,past+q] (cnets_ours.py:43-70, forced fp32 at :791 / :900)."""
    m = torch.full((q_len, q_len), F32_MIN, device=device)
    idx = torch.arange(q_len, device=device)
    m.masked_fill_(idx < (idx + 1).view(q_len, 1), 0)
    if past > 0:
        m = torch.cat([torch.zeros(q_len, past, dtype=torch.float32, device=device), m], dim=-1)
    return m[None, None]


def _mlp(x, w, prefix):
    """SwiGLU (cnets_ours.py:508 / modeling_llama_kv.py:421)."""
    g = F.linear(x, w[prefix + "gate_proj.weight"])
    u = F.linear(x, w[prefix + "up_proj.weight"])
    return F.linear(F.silu(g) * u, w[prefix + "down_proj.weight"])


# --------------------------------------------------------------------------------------
# target LM with static KV slab and tree mask   (A6, A10)
# --------------------------------------------------------------------------------------
class TargetOracle:
    """modeling_llama_kv.LlamaForCausalLM + kv_cache.KVCache restated.

    ``kv`` is the reference slab layout [2*layers, 1, kv_heads, max_pos, D] (kv_cache.py:118-126);
    ``cur_len`` plays the role of ``current_length_data`` (one value, batch 1, kv_cache.py:130-132).
    """

    def __init__(self, cfg: TargetCfg, weights: Dict[str, torch.Tensor], kv_max: Optional[int] = None):
        self.cfg = cfg
        self.w = weights
        any_w = weights["lm_head.weight"]
        self.device, self.dtype = any_w.device, any_w.dtype
        self.cos, self.sin = rope_tables(cfg.head_dim, cfg.max_pos, cfg.rope_theta, self.dtype, self.device)
        self.kv_max = kv_max or cfg.max_pos
        self.kv = torch.zeros(2 * cfg.layers, 1, cfg.kv_heads, self.kv_max, cfg.head_dim,
                              dtype=self.dtype, device=self.device)
        self.cur_len = 0

    def reset(self):
        self.cur_len = 0

    def embed(self, ids):
        return F.embedding(ids, self.w["model.embed_tokens.weight"])

    def forward(self, input_ids=None, inputs_embeds=None, position_ids=None, tree_mask=None):
        """LlamaModel.forward (modeling_llama_kv.py:927-1080) + lm_head (:1196-1197).

        Returns (logits fp32 [1,q,V], hidden [1,q,H] = POST-final-norm, :1062-1066)."""
        cfg, w = self.cfg, self.w
        if inputs_embeds is None:
            inputs_embeds = self.embed(input_ids)
        q_len = inputs_embeds.shape[1]
        past = self.cur_len
        if position_ids is None:
            position_ids = torch.arange(past, past + q_len, device=self.device)
        position_ids = position_ids.long()
        pos3 = None
        if cfg.mrope_section is not None:                # Qwen2.5-VL: [3, q]; text tokens carry the same id on all axes
            pos3 = position_ids.view(3, q_len) if position_ids.numel() == 3 * q_len else \
                position_ids.view(1, q_len).expand(3, q_len)
        else:
            position_ids = position_ids.view(1, q_len)
        sdpa = cfg.attn_impl == "sdpa"
        mdt = self.dtype if sdpa else torch.float32      # SDPA class: mask in the model dtype (:1455-1540)

        # mask: modeling_llama_kv.py:890-924.  q_len==1 -> only the all-zero expanded mask.
        if q_len > 1:
            mask = causal_mask(q_len, past, self.device)
            if sdpa:
                mask = torch.where(mask == 0, 0.0, torch.finfo(mdt).min).to(mdt)
            if tree_mask is not None:
                tl = tree_mask.size(-1)
                mask = mask.clone()
                mask[:, :, -tl:, -tl:][tree_mask.to(self.device) == 0] = mask.min()
        else:
            mask = torch.zeros(1, 1, 1, past + 1, dtype=mdt, device=self.device)

        h = inputs_embeds
        nh, nkv, D = cfg.heads, cfg.kv_heads, cfg.head_dim
        for li in range(cfg.layers):
            p = f"model.layers.{li}."
            res = h
            x = rms_norm(h, w[p + "input_layernorm.weight"], cfg.eps)
            qb, kb, vb = (w.get(p + f"self_attn.{n}_proj.bias") for n in "qkv")
            q = F.linear(x, w[p + "self_attn.q_proj.weight"], qb).view(1, q_len, nh, D).transpose(1, 2)
            k = F.linear(x, w[p + "self_attn.k_proj.weight"], kb).view(1, q_len, nkv, D).transpose(1, 2)
            v = F.linear(x, w[p + "self_attn.v_proj.weight"], vb).view(1, q_len, nkv, D).transpose(1, 2)
            if pos3 is not None:
                q, k = apply_mrope(q, k, self.cos, self.sin, pos3, cfg.mrope_section)
            else:
                q, k = apply_rope(q, k, self.cos, self.sin, position_ids)
            # KVCache.cat (kv_cache.py:52-66)
            self.kv[2 * li, :, :, past:past + q_len] = k
            self.kv[2 * li + 1, :, :, past:past + q_len] = v
            K = self.kv[2 * li, :, :, :past + q_len]
            V = self.kv[2 * li + 1, :, :, :past + q_len]
            if nkv != nh:
                rep = nh // nkv
                K = K[:, :, None].expand(1, nkv, rep, past + q_len, D).reshape(1, nh, past + q_len, D)
                V = V[:, :, None].expand(1, nkv, rep, past + q_len, D).reshape(1, nh, past + q_len, D)
            if sdpa:
                # modeling_qwen2_5_vl_kv.py:1153-1160.  The mask is dropped when SDPA can infer it (:1489-1503,
                # `_ignore_causal_mask_sdpa`): a prefill without past runs with is_causal=True, a single token unmasked.
                if q_len == 1:
                    ao = F.scaled_dot_product_attention(q.contiguous(), K.contiguous(), V.contiguous())
                elif past == 0 and tree_mask is None:
                    ao = F.scaled_dot_product_attention(q.contiguous(), K.contiguous(), V.contiguous(), is_causal=True)
                else:
                    ao = F.scaled_dot_product_attention(q.contiguous(), K.contiguous(), V.contiguous(), attn_mask=mask)
                ao = ao.transpose(1, 2).contiguous().reshape(1, q_len, cfg.hidden)
            else:
                # eager attention, modeling_llama_kv.py:602-623
                aw = torch.matmul(q, K.transpose(2, 3)) / math.sqrt(D)
                aw = aw + mask
                aw = F.softmax(aw, dim=-1, dtype=torch.float32).to(q.dtype)
                ao = torch.matmul(aw, V).transpose(1, 2).contiguous().reshape(1, q_len, cfg.hidden)
            h = res + F.linear(ao, w[p + "self_attn.o_proj.weight"])
            res = h
            x = rms_norm(h, w[p + "post_attention_layernorm.weight"], cfg.eps)
            h = res + _mlp(x, w, p + "mlp.")
        self.cur_len = past + q_len
        h = rms_norm(h, w["model.norm.weight"], cfg.eps)
        logits = F.linear(h, w["lm_head.weight"]).float()
        return logits, h


# --------------------------------------------------------------------------------------
# draft model (A1-A5, A7)
# --------------------------------------------------------------------------------------
class DraftOracle:
    """cnets_ours.Model restated (one decoder layer, no input norm, no final norm)."""

    def __init__(self, cfg: DraftCfg, weights: Dict[str, torch.Tensor]):
        self.cfg = cfg
        self.w = weights
        any_w = weights["fc.weight"]
        self.device, self.dtype = any_w.device, any_w.dtype
        self.cos, self.sin = rope_tables(cfg.head_dim, cfg.max_pos, cfg.rope_theta, self.dtype, self.device)
        self.stable_kv = None          # (K [1,h,C,D], V, real_len)      cnets_ours.py:1108
        self.last_img_hidden = None    # global image feature g [1,H]    cnets_ours.py:759,930
        self.tree_mask = None

    def reset_kv(self):
        """cnets_ours.py:1040-1041."""
        self.stable_kv = None

    # -- vision adaptor: token-compression cross-attention (A1; cnets_ours.py:630-661)
    def img_adaptor(self, img_embeds):
        cfg, w = self.cfg, self.w
        n = img_embeds.shape[1]
        nh, D = cfg.heads, cfg.head_dim
        q = w["imadpt.q"].view(1, cfg.num_q, nh, D).transpose(1, 2)
        kb = w.get("imadpt.k_proj.bias")
        vb = w.get("imadpt.v_proj.bias")
        k = F.linear(img_embeds, w["imadpt.k_proj.weight"], kb).view(1, n, nh, D).transpose(1, 2)
        v = F.linear(img_embeds, w["imadpt.v_proj.weight"], vb).view(1, n, nh, D).transpose(1, 2)
        o = F.scaled_dot_product_attention(q.contiguous(), k.contiguous(), v.contiguous(), is_causal=False)
        o = o.transpose(1, 2).contiguous().reshape(1, cfg.num_q, cfg.hidden)
        return F.linear(o, w["imadpt.o_proj.weight"])

    def _fuse(self, emb, hid):
        """h' = fc(cat(emb, img_fc(cat(hid, g))))   (A2; cnets_ours.py:920-922, 982-988)."""
        w = self.w
        g = self.last_img_hidden.expand_as(hid)
        x = F.linear(torch.cat((hid, g), dim=-1), w["img_fc.weight"], w.get("img_fc.bias"))
        return F.linear(torch.cat((emb, x), dim=-1), w["fc.weight"], w.get("fc.bias"))

    def _layer(self, h, mask, position_ids, past):
        """LlamaDecoderLayer index 0 (cnets_ours.py:545-600) with attention :323-463."""
        cfg, w = self.cfg, self.w
        m = h.shape[1]
        nh, D = cfg.heads, cfg.head_dim
        p = "layers.0."
        qb, kb, vb = (w.get(p + f"self_attn.{n}_proj.bias") for n in "qkv")
        q = F.linear(h, w[p + "self_attn.q_proj.weight"], qb).view(1, m, nh, D).transpose(1, 2)
        k = F.linear(h, w[p + "self_attn.k_proj.weight"], kb).view(1, m, nh, D).transpose(1, 2)
        v = F.linear(h, w[p + "self_attn.v_proj.weight"], vb).view(1, m, nh, D).transpose(1, 2)
        q, k = apply_rope(q, k, self.cos, self.sin, position_ids)
        if past is not None:
            k = torch.cat([past[0], k], dim=2)
            v = torch.cat([past[1], v], dim=2)
        new_past = (k, v, int(position_ids.max()) + 1)     # 3-tuple cache, :416-418
        ao = F.scaled_dot_product_attention(q.contiguous(), k.contiguous(), v.contiguous(),
                                            attn_mask=mask.to(q.dtype))
        ao = ao.transpose(1, 2).contiguous().reshape(1, m, cfg.hidden)
        h = h + F.linear(ao, w[p + "self_attn.o_proj.weight"])
        x = rms_norm(h, w[p + "post_attention_layernorm.weight"], cfg.eps)
        h = h + _mlp(x, w, p + "mlp.")
        return h, new_past

    def pack_prefill(self, hidden, inputs_embeds, image_mask):
        """Image/prefill branch of Model.forward (cnets_ours.py:879-975).

        hidden, inputs_embeds: [1,L,H] (already shifted by the caller, :1082); image_mask [1,L] bool
        in the UNshifted frame.  Returns packed rows [1,n_c,H] and their position ids [1,n_c].
        Exactly one contiguous image run is supported (the reference breaks on more)."""
        cfg = self.cfg
        L = hidden.shape[1]
        pos = torch.arange(L, device=self.device)
        im = image_mask[0, 1:]                                           # :880
        ends = torch.cat([im[:-1] & ~im[1:], im[-1:]])                   # :881-883
        last_ids = torch.where(ends)[0]
        if last_ids.numel() != 1:
            raise NotImplementedError("exactly one contiguous image run is supported")
        self.last_img_hidden = torch.zeros_like(hidden[0, :1])           # :914
        end = int(last_ids[0]) + 1
        cur = im[:end]
        rows, pids = [], []
        rows.append(self._fuse(inputs_embeds[0, :end][~cur], hidden[0, :end][~cur]))      # :918-922
        adapted = self.img_adaptor(inputs_embeds[0, :end][cur].unsqueeze(0)).squeeze(0)   # :924-927
        rows.append(adapted[:-1])                                                          # :928
        self.last_img_hidden = adapted[-1:]                                                # :930
        pids += [pos[:end][~cur], pos[end - adapted.shape[0] + 1:end]]                     # :932-937
        rows.append(self._fuse(inputs_embeds[0, end:], hidden[0, end:]))                   # :944-948
        pids.append(pos[end:])
        return torch.cat(rows, 0).unsqueeze(0), torch.cat(pids, 0).unsqueeze(0)

    def forward(self, hidden, input_ids=None, inputs_embeds=None, past=None, position_ids=None,
                image_mask=None):
        """Model.forward (cnets_ours.py:817-1038) with use_cache=True.  Returns (out [1,m',H], new_past)."""
        w = self.w
        m = hidden.shape[1]
        past_len = 0 if past is None else past[0].shape[2]
        real_len = 0 if past is None else past[2]
        if inputs_embeds is None:
            inputs_embeds = F.embedding(input_ids, w["embed_tokens.weight"])
        inputs_embeds = inputs_embeds.to(hidden)
        if position_ids is None:
            position_ids = torch.arange(real_len, real_len + m, device=self.device).unsqueeze(0)
        else:
            position_ids = position_ids.view(1, m).long()

        if image_mask is not None and past is None:
            h, position_ids = self.pack_prefill(hidden, inputs_embeds, image_mask)
            mask = causal_mask(h.shape[1], 0, self.device)                # :971-975
        else:
            # mask: _prepare_decoder_attention_mask, cnets_ours.py:781-815
            if m > 1:
                mask = causal_mask(m, past_len, self.device)
            else:
                mask = torch.zeros(1, 1, 1, past_len + 1, dtype=torch.float32, device=self.device)
            if self.tree_mask is not None:
                t0, t1 = self.tree_mask.shape[-2:]
                mask = mask.clone()
                mask[:, :, -t0:, -t1:][self.tree_mask == 0] = F32_MIN
            if past is None:
                self.last_img_hidden = torch.zeros_like(hidden[0, :1])    # :977-981 (dummy adaptor call adds 0)
            h = self._fuse(inputs_embeds[0], hidden[0]).unsqueeze(0)
        return self._layer(h, mask, position_ids, past)

    # -- draft-tree expansion + tree build (cnets_ours.py:1044-1238)
    @torch.no_grad()
    def topk_generate(self, hidden, input_ids, head_weight, inputs_embeds=None, image_mask=None,
                      trace: Optional[dict] = None):
        cfg, w = self.cfg, self.w
        top_k, depth, total = cfg.top_k, cfg.depth, cfg.total_tokens - 1
        sample_token = input_ids[:, -1]
        if inputs_embeds is not None:                                     # :1066-1082
            new = F.embedding(input_ids[:, inputs_embeds.shape[-2]:], w["embed_tokens.weight"])
            inputs_embeds = torch.cat((inputs_embeds[:, 1:], new), dim=-2)
        input_ids = input_ids[:, 1:]
        len_posi = input_ids.shape[1]
        self.tree_mask = None
        if self.stable_kv is not None:                                    # :1090-1097
            out, past = self.forward(hidden, input_ids=input_ids[:, -hidden.shape[1]:], past=self.stable_kv)
        elif inputs_embeds is not None:
            out, past = self.forward(hidden, inputs_embeds=inputs_embeds, image_mask=image_mask)
        else:
            out, past = self.forward(hidden, input_ids=input_ids)
        self.stable_kv = past
        last_hidden = out[:, -1]
        root_logits = F.linear(last_hidden, head_weight)
        last_p = F.log_softmax(root_logits, dim=-1)                          # bf16, :1111-1113
        top = torch.topk(last_p, top_k, dim=-1)
        scores = top.values[0]
        scores_list = [scores[None]]
        parents_list = [torch.zeros(1, dtype=torch.long, device=self.device)]
        ss_token = [top.indices]
        ids = top.indices
        in_hidden = last_hidden[None].repeat(1, top_k, 1)
        eye = torch.eye(top_k, device=self.device)[None, None]
        tree_mask = eye
        cs_index = torch.arange(top_k, device=self.device)
        if trace is not None:
            trace["stable_out"] = last_hidden.clone()
            trace["levels"] = []
            trace["logit_absmax"] = [float(root_logits.float().abs().max())]   # noise unit of the margin rule (tests/parity.py)
        for i in range(depth):                                            # :1126-1165
            self.tree_mask = tree_mask
            pos = len_posi + torch.zeros(top_k, dtype=torch.long, device=self.device)
            out, past = self.forward(in_hidden, input_ids=ids, past=past, position_ids=pos)
            len_posi += 1
            bias = 1 + top_k ** 2 * max(0, i - 1) + (top_k if i > 0 else 0)
            parents_list.append(cs_index + bias)
            level_logits = F.linear(out[0], head_weight)
            last_p = F.log_softmax(level_logits, dim=-1)
            top = torch.topk(last_p, top_k, dim=-1)
            cu_scores = top.values + scores[:, None]
            cs = torch.topk(cu_scores.view(-1), top_k, dim=-1)
            cs_index, scores = cs.indices, cs.values
            out_ids = cs_index // top_k
            in_hidden = out[:, out_ids]
            ids = top.indices.view(-1)[cs_index][None]
            ss_token.append(top.indices)
            scores_list.append(cu_scores)
            tree_mask = torch.cat((tree_mask[:, :, out_ids], eye), dim=3)
            if trace is not None:
                trace["levels"].append(dict(logp=last_p.clone(), topk_idx=top.indices.clone(),
                                            cu_scores=cu_scores.clone(), cs_index=cs_index.clone()))
                trace["logit_absmax"].append(float(level_logits.float().abs().max()))
        self.tree_mask = None
        scores_flat = torch.cat(scores_list, dim=0).view(-1)
        tokens_flat = torch.cat(ss_token, dim=0).view(-1)
        parents = torch.cat(parents_list, dim=0)
        if trace is not None:
            trace.update(scores_flat=scores_flat.clone(), tokens_flat=tokens_flat.clone(),
                         parents=parents.clone())
        return tree_build(scores_flat, tokens_flat, parents, sample_token, total, top_k)


def tree_build(scores_flat, tokens_flat, parents, sample_token, total: int, top_k: int, sort_paths=False):
    """Tree build (A7; cnets_ours.py:1167-1238).

    scores_flat [top_k + depth*top_k^2], tokens_flat same length, parents [1 + depth*top_k] (node id
    of the parent of each group of top_k siblings).  Returns draft_tokens [1,total+1] i64,
    retrieve_indices [n_leaves, max_depth] i64 (pad -1), tree_mask [1,1,total+1,total+1] f32,
    tree_position_ids [total+1] i64."""
    top_idx = torch.topk(scores_flat, total, dim=-1).indices
    top_idx = torch.sort(top_idx).values
    draft_tokens = torch.cat((sample_token, tokens_flat[top_idx]), dim=0)
    draft_parents = parents[top_idx // top_k].long()
    mask_index = torch.searchsorted(top_idx, draft_parents - 1, right=False)
    mask_index[draft_parents == 0] = -1
    mask_index = (mask_index + 1).tolist()          # parent slot (0 = root) of node i+1
    n = total + 1
    tm = torch.eye(n).bool()
    tm[:, 0] = True
    for i in range(total):
        tm[i + 1] |= tm[mask_index[i]]
    pos = tm.sum(dim=1) - 1
    max_depth = int(pos.max()) + 1
    noleaf = set(mask_index)
    paths = []
    pos_l = pos.tolist()
    for i in range(n):
        if i in noleaf:
            continue
        path = [-1] * max_depth
        cid = i
        for j in range(pos_l[i], -1, -1):
            path[j] = cid
            cid = mask_index[cid - 1]
        paths.append(path)
    if sort_paths:                                   # T>0 only, cnets_ours.py:1215-1224
        big = total + 5
        paths = sorted(paths, key=lambda p: [x if x >= 0 else big for x in p])
    retrieve = torch.tensor(paths, dtype=torch.long)
    return draft_tokens[None], retrieve, tm.float()[None, None], pos.to(scores_flat.device)


# --------------------------------------------------------------------------------------
# accept + splice (A8, A9)
# --------------------------------------------------------------------------------------
def evaluate_posterior_greedy(logits, candidates):
    """utils.py:438-451.  logits [n_leaves, depth, V] fp32; candidates [n_leaves, depth] (pad -1)."""
    post = (candidates[:, 1:].to(logits.device) == torch.argmax(logits[:, :-1], dim=-1)).int()
    acc = torch.cumprod(post, dim=1).sum(dim=1)
    accept_length = int(acc.max())
    best = 0 if accept_length == 0 else int(torch.argmax(acc))
    return best, accept_length, logits[best, accept_length]


def warp_logits(logits, temperature: float, top_p: float = 0.0, top_k: int = 0):
    """The logits processors `prepare_logits_processor` assembles for T > 0 (utils.py:39-55), applied in its order:
    temperature (only when != 1), nucleus filter for 1e-8 <= top_p < 1, top-k filter for top_k > 0.  The warpers are
    third-party code (transformers.generation.logits_process, pinned 4.49.0 in requirements.txt:219; 5.5 installed here),
    restated from their published algorithms:
      * TemperatureLogitsWarper: logits / temperature;
      * TopPLogitsWarper: sort ascending, softmax, running sum; drop every token whose running sum is <= 1 - top_p, but
        never the most probable one (min_tokens_to_keep = 1);
      * TopKLogitsWarper: drop every token whose logit is below the k-th largest.
    Dropped tokens get -inf.  `logits` [rows, V] fp32; returns a new tensor.  (The repetition penalty sits between the
    temperature and the nucleus filter in the reference; no evaluation script sets it and it is not restated.)
    Pinned against the reference's own processor list in tests/test_oracle_golden.py (tests/golden/warpers_sampling.pt)."""
    x = logits.clone()
    if temperature > 1e-5 and temperature != 1.0:
        x = x / temperature
    if 1e-8 <= top_p < 1.0:
        asc, idx = torch.sort(x, dim=-1)                       # ascending
        running = torch.cumsum(torch.softmax(asc, dim=-1), dim=-1)
        drop_sorted = running <= (1.0 - top_p)
        drop_sorted[..., -1] = False
        drop = torch.zeros_like(drop_sorted).scatter(-1, idx, drop_sorted)
        x = x.masked_fill(drop, float("-inf"))
    if top_k > 0:
        k = min(int(top_k), x.shape[-1])
        kth = torch.topk(x, k, dim=-1).values[..., -1:]
        x = x.masked_fill(x < kth, float("-inf"))
    return x


def evaluate_posterior_sampling(logits, candidates, temperature: float, uni, top_p: float = 0.0, top_k: int = 0):
    """utils.py:453-493 (T > 0): SpecInfer-style acceptance along the tree with residual renormalisation.

    logits [n_leaves, depth, V] fp32 (tree logits gathered along every root->leaf path, utils.py:411);
    candidates [n_leaves, depth] (pad -1); `uni` [n_leaves, depth] are the uniforms the reference draws with
    torch.rand_like (:456) -- passed in so that the decisions can be reproduced.  The logits processor is the
    TemperatureLogitsWarper of prepare_logits_processor (:39-55): logits / temperature.
    Returns (best_candidate, accept_length, sample_p [V]) like the reference (accept_length already minus one)."""
    if top_p or top_k:                                 # the processor list works on [rows, V] (utils.py:454)
        V = logits.shape[-1]
        logits_p = torch.softmax(warp_logits(logits.reshape(-1, V), temperature, top_p, top_k), dim=-1).view_as(logits)
    else:
        logits_p = torch.softmax(logits / temperature, dim=-1)
    accept_length = 1
    accept_cand = candidates[0].clone()
    best_candidate = 0
    adjustflag = False
    gtp = None
    for i in range(1, candidates.shape[1]):
        if i != accept_length:
            break
        adjustflag = False
        is_eq = (candidates[:, :accept_length] == accept_cand[:accept_length]).all(dim=1)
        fi = int(torch.nonzero(is_eq)[0, 0])
        gtp = logits_p[fi, i - 1].clone()
        seen = set()
        for j in range(candidates.shape[0]):
            if not bool(is_eq[j]):
                continue
            xi = int(candidates[j, i])
            if xi == -1 or xi in seen:
                continue
            seen.add(xi)
            if float(uni[j, i]) <= float(gtp[xi]):
                accept_cand[accept_length] = xi
                accept_length += 1
                best_candidate = j
                break
            gtp[xi] = 0
            gtp = gtp / gtp.sum()
            adjustflag = True
    if adjustflag and accept_length != candidates.shape[1]:
        sample_p = gtp
    else:
        sample_p = logits_p[best_candidate, accept_length - 1]
    return best_candidate, accept_length - 1, sample_p


def kv_splice(kv, select_indices, prev_len: int):
    """KV-cache splice (utils.py:529-541): rows select_indices -> prev_len.. in every (layer,K/V,head)."""
    tgt = kv[..., select_indices.to(kv.device), :]
    kv[..., prev_len:prev_len + tgt.shape[-2], :].copy_(tgt)
    return prev_len + tgt.shape[-2]


# --------------------------------------------------------------------------------------
# the outer loop (spec_model_ours.py:455-547) and the plain greedy loop it must equal
# --------------------------------------------------------------------------------------
@dataclass
class SpecTrace:
    cycles: List[dict] = field(default_factory=list)
    draft: List[dict] = field(default_factory=list)   # per topk_generate call (prefill first): scores / margins of the tree


@torch.no_grad()
def spec_generate(target: TargetOracle, draft: DraftOracle, input_ids, inputs_embeds, image_mask,
                  max_new_tokens: int, eos_token_id: Optional[int] = None, max_cycles: int = 2048,
                  trace: Optional[SpecTrace] = None):
    """specgenerate at T=0 (spec_model_ours.py:248-582) with initialize_tree (utils.py:266-327),
    tree_decoding (:389-412), evaluate_posterior (:438-451), update_inference_inputs (:497-593).

    Returns (output_ids [1,L+n], new_token, idx, acceptance_len list)."""
    head_w = target.w["lm_head.weight"]
    draft.reset_kv()
    target.reset()
    input_ids = input_ids.clone()
    input_len = input_ids.shape[1]
    if inputs_embeds is None:
        logits, hidden = target.forward(input_ids=input_ids)
    else:
        logits, hidden = target.forward(inputs_embeds=inputs_embeds)
    token = torch.argmax(logits[:, -1])[None, None]
    input_ids = torch.cat((input_ids, token), dim=1)
    dt = {} if trace is not None else None
    draft_tokens, retrieve, tree_mask, tree_pos = draft.topk_generate(
        hidden, input_ids, head_w, inputs_embeds=inputs_embeds, image_mask=image_mask, trace=dt)
    if trace is not None:
        trace.draft.append(dt)
    input_ids = input_ids[:, :-1]     # initialize_tree's cat is local (utils.py:292); caller keeps the prompt
    new_token, accept_lens, idx = 0, [], 0
    pad = torch.full((1, 1), -1, dtype=torch.long, device=input_ids.device)
    for idx in range(max_cycles):
        prev = input_ids.shape[1]
        t_logits, t_hidden = target.forward(input_ids=draft_tokens, position_ids=tree_pos + prev,
                                            tree_mask=tree_mask)
        logits = t_logits[0, retrieve]
        cand = torch.cat((draft_tokens, pad), dim=1)[0, retrieve]
        best, a, sample_p = evaluate_posterior_greedy(logits, cand)
        accept_lens.append(a)
        sel = retrieve[best, :a + 1] + prev
        input_ids = torch.cat([input_ids, cand[None, best, :a + 1]], dim=-1)
        target.cur_len = kv_splice(target.kv, sel, prev)
        acc_hidden = t_hidden[:, retrieve][:, best, :a + 1]
        token = torch.argmax(sample_p)[None, None]
        if trace is not None:
            trace.cycles.append(dict(draft_tokens=draft_tokens.clone(), retrieve=retrieve.clone(),
                                     tree_mask=tree_mask.clone(), tree_pos=tree_pos.clone(),
                                     best=best, accept=a, token=int(token), logits_root=t_logits[0, 0].clone()))
        dt = {} if trace is not None else None
        draft_tokens, retrieve, tree_mask, tree_pos = draft.topk_generate(
            acc_hidden, torch.cat((input_ids, token), dim=1), head_w, trace=dt)
        if trace is not None:
            trace.draft.append(dt)
        new_token += a + 1
        if eos_token_id is not None and eos_token_id in input_ids[0, input_len:].tolist():
            break
        if new_token > max_new_tokens:
            break
    return input_ids, new_token, idx, accept_lens


@torch.no_grad()
def greedy_generate(target: TargetOracle, input_ids, inputs_embeds, n_tokens: int):
    """Plain autoregressive greedy decode over the same KV slab (the loop of
    vispec/evaluation/gen_baseline_answer_coco_caption.py:34-133).  spec_generate must equal it."""
    target.reset()
    if inputs_embeds is None:
        logits, _ = target.forward(input_ids=input_ids)
    else:
        logits, _ = target.forward(inputs_embeds=inputs_embeds)
    out = input_ids.clone()
    for _ in range(n_tokens):
        tok = torch.argmax(logits[:, -1])[None, None]
        out = torch.cat((out, tok), dim=1)
        logits, _ = target.forward(input_ids=tok)
    return out
