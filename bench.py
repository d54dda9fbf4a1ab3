#!/usr/bin/env python
"""bench.py -- decode tokens/s (+ accept length) of the ViSpec draft-and-verify path on B200.

Workload (BASELINE.json configs[1]): LLaVA-1.6-Vicuna-7B shapes (32 layers, H=4096, 32x128 heads, I=11008,
V=32064) + ViSpec draft head (1 layer, total_token=30, depth=3, top_k=8, num_q=2), T=0 greedy, synthetic
prompt of 576 image-embedding rows + 128 text tokens, 256 generated tokens, random-init bf16 weights (no
checkpoints offline).  One "step" = one generation of 256 tokens.

  value : decode tokens/s with the prompt already resident in HBM, prefill excluded (the reference's
          `return_decode_time` region, spec_model_ours.py:480-482,575-578), CUDA events, max over ranks.
  e2e   : the same metric through the public API `SpecModel.specgenerate` with HOST (pinned) inputs: H2D copy of
          ids/embeddings/mask, prefill, decode and the D2H read of the output ids are all inside the timed region.
  roofline     : the dominant kernel = the persistent program kernel (ONE launch per draft+verify cycle): algorithmic
                 bytes per launch (every weight matrix of the 32-layer verify step and of the 4 draft forwards read once
                 + the KV rows read once) / the kernel's average launch duration over the timed region
                 (CUDA events on the launching stream), against the measured HBM copy bandwidth.
  roofline_attention : the north star's KV-read figure: K+V bytes of one layer / the duration of that layer's two
                 attention phases inside the kernel (per-phase %globaltimer trace of the same kernel).
  cpu_baseline : the oracle (CPU restatement of the reference) timed on this box's host cores on a bounded sample.

`--impl reference` times the reference's CPU implementation on the host cores (the unmodified vispec.model modules
vendored under baseline/_ref; the oracle port only if that copy is missing); `--impl reference-gpu` and the
`gpu_eager_baseline` object time the same modules with `.cuda()` on this B200 (the north star's denominator).
`--gpus N` (under torchrun): N independent replicas, one prompt stream each, no data-path collective (weak scaling).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# dram__bytes_read.sum + dram__bytes_write.sum of one program_kernel launch from the committed `ncu --set full` capture
# (profiles/README.md); None until a capture of the current kernel exists
NCU_TRAFFIC_BYTES = 16.810e9 + 0.120e9        # dram__bytes_read.sum + dram__bytes_write.sum of one launch (ncu --set full)
NCU_TRAFFIC_SOURCE = "profiles/r2_ncu_full_program_kernel.csv"
N_PRE, N_IMG, N_POST = 5, 576, 123          # 576 image + 128 text tokens = 704
GEN_TOKENS = 256
WORKLOAD = "LLaVA-1.6-Vicuna-7B shapes + ViSpec head, T=0, 576 img + 128 text tokens, 256-token gen, one stream per B200"


# ------------------------------------------------------------------------------------------------------------
# synthetic model (random-init weights of the named architecture, generated on the device)
# ------------------------------------------------------------------------------------------------------------
def _randn(gen, *shape, std, device):
    return (torch.randn(*shape, generator=gen, device=device, dtype=torch.float32) * std).to(torch.bfloat16)


def synth_target_weights(cfg, seed, device, gain=0.5):
    g = torch.Generator(device=device).manual_seed(seed)
    H, I, V, D = cfg.hidden, cfg.inter, cfg.vocab, 128
    w = {"model.embed_tokens.weight": _randn(g, V, H, std=1.0, device=device)}
    for li in range(cfg.layers):
        p = f"model.layers.{li}."
        w[p + "self_attn.q_proj.weight"] = _randn(g, cfg.heads * D, H, std=H ** -0.5, device=device)
        w[p + "self_attn.k_proj.weight"] = _randn(g, cfg.kv_heads * D, H, std=H ** -0.5, device=device)
        w[p + "self_attn.v_proj.weight"] = _randn(g, cfg.kv_heads * D, H, std=H ** -0.5, device=device)
        w[p + "self_attn.o_proj.weight"] = _randn(g, H, H, std=gain * H ** -0.5, device=device)
        w[p + "mlp.gate_proj.weight"] = _randn(g, I, H, std=H ** -0.5, device=device)
        w[p + "mlp.up_proj.weight"] = _randn(g, I, H, std=H ** -0.5, device=device)
        w[p + "mlp.down_proj.weight"] = _randn(g, H, I, std=gain * I ** -0.5, device=device)
        w[p + "input_layernorm.weight"] = (1.0 + 0.1 * torch.randn(H, generator=g, device=device)).to(torch.bfloat16)
        w[p + "post_attention_layernorm.weight"] = (1.0 + 0.1 * torch.randn(H, generator=g, device=device)).to(torch.bfloat16)
    w["model.norm.weight"] = (1.0 + 0.1 * torch.randn(H, generator=g, device=device)).to(torch.bfloat16)
    w["lm_head.weight"] = _randn(g, V, H, std=4.0 * H ** -0.5, device=device)
    return w


def synth_draft_weights(cfg, seed, device, gain=0.5):
    g = torch.Generator(device=device).manual_seed(seed)
    H, I, V, D = cfg.hidden, cfg.inter, cfg.vocab, 128
    w = {"embed_tokens.weight": _randn(g, V, H, std=1.0, device=device)}
    p = "layers.0."
    for n in "qkv":
        w[p + f"self_attn.{n}_proj.weight"] = _randn(g, H, H, std=H ** -0.5, device=device)
    w[p + "self_attn.o_proj.weight"] = _randn(g, H, H, std=gain * H ** -0.5, device=device)
    w[p + "mlp.gate_proj.weight"] = _randn(g, I, H, std=H ** -0.5, device=device)
    w[p + "mlp.up_proj.weight"] = _randn(g, I, H, std=H ** -0.5, device=device)
    w[p + "mlp.down_proj.weight"] = _randn(g, H, I, std=gain * I ** -0.5, device=device)
    w[p + "post_attention_layernorm.weight"] = (1.0 + 0.1 * torch.randn(H, generator=g, device=device)).to(torch.bfloat16)
    for n in ("fc", "img_fc"):
        w[n + ".weight"] = _randn(g, H, 2 * H, std=(2 * H) ** -0.5, device=device)
        if cfg.bias:
            w[n + ".bias"] = _randn(g, H, std=0.1, device=device)
    w["imadpt.q"] = _randn(g, cfg.num_q, cfg.heads, D, std=D ** -0.5, device=device)
    for n in ("k_proj", "v_proj", "o_proj"):
        w[f"imadpt.{n}.weight"] = _randn(g, H, H, std=H ** -0.5, device=device)
    return w


def synth_prompt(cfg, embed, seed, device, n_post=None):
    """5 text + 576 image + `n_post` text tokens (default: the 128-text-token workload of BASELINE config 2; config 5
    draws 64-512 text tokens per stream)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    L = N_PRE + N_IMG + (N_POST if n_post is None else n_post)
    ids = torch.randint(1, cfg.vocab - 64, (1, L), generator=g)
    mask = torch.zeros(1, L, dtype=torch.bool)
    mask[0, N_PRE:N_PRE + N_IMG] = True
    ids[mask] = cfg.vocab - 1                                   # image placeholder id
    emb = torch.nn.functional.embedding(ids.to(device), embed).clone()
    img = (torch.randn(N_IMG, cfg.hidden, generator=g) * 0.02).to(torch.bfloat16)
    emb[0, N_PRE:N_PRE + N_IMG] = img.to(device)
    return ids.to(device), emb, mask.to(device)


def make_bigram(tw, dw, tcfg, dcfg, fc_noise=0.0, seed=0):
    """The tau > 0 fixture of SURVEY.md 7.2 at full size (oracle/synth.py:make_bigram_fixture is the CPU original):
    every target layer's o_proj / down_proj is zero, so next = argmax(lm_head(norm(embed[tok]))), and the draft
    (same embedding, fc = [I | 0], zero o_proj / down_proj) predicts it: accept length = depth + 1."""
    for li in range(tcfg.layers):
        tw[f"model.layers.{li}.self_attn.o_proj.weight"].zero_()
        tw[f"model.layers.{li}.mlp.down_proj.weight"].zero_()
    tw["model.norm.weight"].fill_(1.0)
    dw["embed_tokens.weight"] = tw["model.embed_tokens.weight"]
    H = dcfg.hidden
    fc = torch.zeros(H, 2 * H, dtype=torch.bfloat16, device=dw["fc.weight"].device)
    fc[:, :H] = torch.eye(H, dtype=torch.bfloat16, device=fc.device)
    if fc_noise > 0:        # an imperfect draft: intermediate accept lengths (oracle/synth.py:make_bigram_fixture)
        g = torch.Generator(device=fc.device).manual_seed(seed + 2)
        fc = (fc.float() + fc_noise * torch.randn(H, 2 * H, generator=g, device=fc.device) / math.sqrt(2 * H)).to(torch.bfloat16)
    dw["fc.weight"] = fc
    if "fc.bias" in dw:
        dw["fc.bias"].zero_()
    dw["layers.0.self_attn.o_proj.weight"].zero_()
    dw["layers.0.mlp.down_proj.weight"].zero_()


def build_llava7b(device="cuda", seed=0, layers=32, fixture="random", fc_noise=0.0, n_post=None):
    """(model, input_ids, inputs_embeds, image_mask) for the BASELINE config-2 workload.  fixture="bigram": the same
    shapes with weights under which the draft is always right (accept length 4), to exercise accept / splice / the
    accepted-rows draft forward at full size."""
    from vispec_b200.engine import DraftConfig, TargetConfig
    from vispec_b200.model.spec_model_ours import SpecModel

    tcfg = TargetConfig(layers=layers)
    dcfg = DraftConfig()
    tw = synth_target_weights(tcfg, seed, device)
    dw = synth_draft_weights(dcfg, seed + 1, device)
    if fixture == "bigram":
        make_bigram(tw, dw, tcfg, dcfg, fc_noise=fc_noise, seed=seed)
    model = SpecModel.from_state_dicts(tcfg, tw, dcfg, dw, device=device, eos_token_id=None)
    ids, emb, mask = synth_prompt(tcfg, model.base_model_engine.embed, seed + 7, device, n_post=n_post)
    del tw, dw
    torch.cuda.empty_cache()
    return model, ids, emb, mask


def build_qwen7b(device="cuda", seed=0):
    """BASELINE config-3 shapes (Qwen2.5-VL-7B target: 28 q / 4 kv heads, q/k/v biases, theta = 1e6, SDPA semantics,
    multimodal RoPE over a 24 x 24 image grid in the prompt; MHA draft with the same width)."""
    from vispec_b200.engine import DraftConfig, TargetConfig
    from vispec_b200.model import utils as vu
    from vispec_b200.model.spec_model_ours import SpecModel

    tcfg = TargetConfig(hidden=3584, inter=18944, layers=28, heads=28, kv_heads=4, vocab=152064, max_pos=8192, eps=1e-6,
                        rope_theta=1e6, mrope_section=(16, 24, 24), attn_impl="sdpa")
    dcfg = DraftConfig(hidden=3584, inter=18944, heads=28, vocab=152064, max_pos=4096, eps=1e-6, rope_theta=1e6)
    tw = synth_target_weights(tcfg, seed, device)
    g = torch.Generator(device=device).manual_seed(seed + 2)
    for li in range(tcfg.layers):
        for n, rows in (("q", tcfg.heads * 128), ("k", tcfg.kv_heads * 128), ("v", tcfg.kv_heads * 128)):
            tw[f"model.layers.{li}.self_attn.{n}_proj.bias"] = _randn(g, rows, std=0.3, device=device)
    dw = synth_draft_weights(dcfg, seed + 1, device)
    model = SpecModel.from_state_dicts(tcfg, tw, dcfg, dw, device=device, eos_token_id=None)
    del tw, dw
    torch.cuda.empty_cache()
    n_pre, grid, n_post, img_id = 5, 24, 123, 151655
    cg = torch.Generator().manual_seed(seed + 7)
    ids = torch.randint(1, 150000, (1, n_pre + grid * grid + n_post), generator=cg)
    ids[0, n_pre:n_pre + grid * grid] = img_id
    pos3, _ = vu.qwen_rope_index(ids, image_grid_thw=[[1, 2 * grid, 2 * grid]], image_token_id=img_id)
    emb = model.base_model_engine.embed_tokens(ids.to(device)).clone()
    emb[0, n_pre:n_pre + grid * grid] = (torch.randn(grid * grid, tcfg.hidden, generator=cg) * 0.02).to(torch.bfloat16).to(device)
    return model, ids.to(device), emb, (ids == img_id).to(device), pos3.to(device), tcfg, dcfg


# ------------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.path = tempfile.mktemp(prefix="vs_clocks_", suffix=".csv")
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for n, v in zip(names, f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            os.unlink(self.path)
        except OSError:
            pass
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


# ------------------------------------------------------------------------------------------------------------
# measured peaks
# ------------------------------------------------------------------------------------------------------------
def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            d = json.load(f)
        for k in ("hbm_gbs", "hbm_gbs_burst", "hbm_gb_s"):
            if k in d:
                return float(d[k]), "measured (MEASURED_PEAKS.json)"
    except (OSError, ValueError):
        pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------------------
# CPU baseline (oracle port) -- the only place bench.py executes oracle/
# ------------------------------------------------------------------------------------------------------------
def _cpu_7b_weights(layers: int):
    """Random bf16 weights at the BASELINE config-2 shapes on the host.  One layer is drawn and cloned per layer
    (distinct memory, same values: timing-equivalent, and 32 x fewer random numbers to draw on the CPU)."""
    from oracle import vispec_oracle as vo

    tcfg, dcfg = vo.TargetCfg(layers=layers, max_pos=2048), vo.DraftCfg()
    g = torch.Generator().manual_seed(0)
    H, I, V = tcfg.hidden, tcfg.inter, tcfg.vocab

    def rn(*s, std):
        return (torch.randn(*s, generator=g) * std).to(torch.bfloat16)

    proto = {"self_attn.q_proj.weight": rn(H, H, std=H ** -0.5), "self_attn.k_proj.weight": rn(H, H, std=H ** -0.5),
             "self_attn.v_proj.weight": rn(H, H, std=H ** -0.5), "self_attn.o_proj.weight": rn(H, H, std=0.5 * H ** -0.5),
             "mlp.gate_proj.weight": rn(I, H, std=H ** -0.5), "mlp.up_proj.weight": rn(I, H, std=H ** -0.5),
             "mlp.down_proj.weight": rn(H, I, std=0.5 * I ** -0.5),
             "input_layernorm.weight": torch.ones(H, dtype=torch.bfloat16),
             "post_attention_layernorm.weight": torch.ones(H, dtype=torch.bfloat16)}
    tw = {"model.embed_tokens.weight": rn(V, H, std=1.0), "model.norm.weight": torch.ones(H, dtype=torch.bfloat16),
          "lm_head.weight": rn(V, H, std=4.0 * H ** -0.5)}
    for li in range(tcfg.layers):
        for k, v in proto.items():
            tw[f"model.layers.{li}.{k}"] = v.clone() if li else v
    dw = {"embed_tokens.weight": tw["model.embed_tokens.weight"]}
    for k, v in proto.items():
        if k != "input_layernorm.weight":
            dw["layers.0." + k] = v
    for n in ("fc", "img_fc"):
        dw[n + ".weight"] = rn(H, 2 * H, std=(2 * H) ** -0.5)
        dw[n + ".bias"] = rn(H, std=0.1)
    dw["imadpt.q"] = rn(dcfg.num_q, dcfg.heads, 128, std=128 ** -0.5)
    for n in ("k_proj", "v_proj", "o_proj"):
        dw[f"imadpt.{n}.weight"] = rn(H, H, std=H ** -0.5)
    L = N_PRE + N_IMG + N_POST
    ids = torch.randint(1, V - 64, (1, L), generator=g)
    mask = torch.zeros(1, L, dtype=torch.bool)
    mask[0, N_PRE:N_PRE + N_IMG] = True
    emb = torch.nn.functional.embedding(ids, tw["model.embed_tokens.weight"])
    return tcfg, dcfg, tw, dw, ids, emb, mask, g


def cpu_reference_sample(n_cycles: int, warm: int, threads: int, layers: int = 32):
    """Time `n_cycles` draft+verify cycles of the reference's CPU path on the host cores at the BASELINE config-2 shapes,
    after `warm` untimed cycles.  Preferred: the UNMODIFIED reference modules vendored under baseline/_ref
    (baseline/ref_harness.py, kind "reference"); fallback: the oracle restatement (kind "port").
    Decode only: a CPU prefill of 704 tokens through 7B weights takes minutes, so the KV state is synthetic -- target KV
    rows [0,704) hold random bf16 values and the draft is prefilled from random hidden states; the per-cycle work
    (13.2 GB of target weights, 3.2 GB of draft weights, the attention over S~734 keys, tree build / accept / splice) is
    exactly that of a real cycle.  Returns (tokens, seconds, accept lengths, kind, what)."""
    torch.set_num_threads(threads)
    tcfg, dcfg, tw, dw, ids, emb, mask, g = _cpu_7b_weights(layers)
    from baseline import ref_harness as rh

    why = rh.available()
    if why is None:
        model = rh.RefSpecStandIn(rh.build_target(tcfg, tw, "cpu", assign=True), rh.build_draft(dcfg, dw, "cpu", assign=True))
        kv = rh.make_kv(model, tcfg, "cpu")
        if warm > 0:
            rh.generate(model, kv, ids, emb, mask, max_new_tokens=10 ** 9, max_cycles=warm, synthetic_prefill=True)
        r = rh.generate(model, kv, ids, emb, mask, max_new_tokens=10 ** 9, max_cycles=n_cycles, synthetic_prefill=True)
        return r["new_token"], r["decode_s"], r["accept"], "reference", \
            "unmodified vispec.model modules (baseline/_ref) driven by baseline/ref_harness.py"
    from oracle import vispec_oracle as vo

    H = tcfg.hidden

    def rn(*s, std):
        return (torch.randn(*s, generator=g) * std).to(torch.bfloat16)

    target, draft = vo.TargetOracle(tcfg, tw, kv_max=1024), vo.DraftOracle(dcfg, dw)
    L = ids.shape[1]
    target.kv[..., :L, :] = rn(2 * tcfg.layers, 1, tcfg.kv_heads, L, 128, std=1.0)
    target.cur_len = L
    ids1 = torch.cat((ids, torch.randint(1, tcfg.vocab - 64, (1, 1), generator=g)), dim=1)
    hidden = rn(1, L, H, std=1.0)
    head_w = tw["lm_head.weight"]
    with torch.no_grad():
        tree = draft.topk_generate(hidden, ids1, head_w, inputs_embeds=emb, image_mask=mask)
        input_ids = ids
        pad = torch.full((1, 1), -1, dtype=torch.long)
        tokens, acc, t0 = 0, [], None
        for c in range(warm + n_cycles):
            if c == warm:
                t0 = time.perf_counter()
                tokens = 0
            draft_tokens, retrieve, tree_mask, tree_pos = tree
            prev = input_ids.shape[1]
            t_logits, t_hidden = target.forward(input_ids=draft_tokens, position_ids=tree_pos + prev, tree_mask=tree_mask)
            logits = t_logits[0, retrieve]
            cand = torch.cat((draft_tokens, pad), dim=1)[0, retrieve]
            best, a, sample_p = vo.evaluate_posterior_greedy(logits, cand)
            sel = retrieve[best, :a + 1] + prev
            input_ids = torch.cat([input_ids, cand[None, best, :a + 1]], dim=-1)
            target.cur_len = vo.kv_splice(target.kv, sel, prev)
            acc_hidden = t_hidden[:, retrieve][:, best, :a + 1]
            token = torch.argmax(sample_p)[None, None]
            tree = draft.topk_generate(acc_hidden, torch.cat((input_ids, token), dim=1), head_w)
            tokens += a + 1
            if c >= warm:
                acc.append(a)
        dt = time.perf_counter() - t0
    return tokens, dt, acc, "port", f"oracle/vispec_oracle.py (vendored reference unavailable: {why})"


def gpu_reference_run(device, steps: int, warmup: int, seed: int, layers: int = 32, count_launches: bool = True):
    """The north star's denominator: the reference's own HF/PyTorch eager path on this B200 -- the unmodified
    vispec.model modules (baseline/_ref) with `.cuda()`, same synthetic 7B weights, same 576 + 128 prompt, 256-token
    greedy generation, prompt pass included in each generation but excluded from the timed region exactly like the
    reference's `return_decode_time` (spec_model_ours.py:480-482,575-578; torch.cuda.synchronize on both sides).
    Returns None when the vendored reference is unavailable."""
    from baseline import ref_harness as rh
    from vispec_b200.engine import DraftConfig, TargetConfig

    why = rh.available()
    if why is not None:
        return {"unavailable": why}
    tcfg, dcfg = TargetConfig(layers=layers), DraftConfig()
    tw = synth_target_weights(tcfg, seed, device)
    dw = synth_draft_weights(dcfg, seed + 1, device)
    model = rh.RefSpecStandIn(rh.build_target(tcfg, tw, device, assign=True), rh.build_draft(dcfg, dw, device, assign=True))
    ids, emb, mask = synth_prompt(tcfg, tw["model.embed_tokens.weight"], seed + 7, device)
    kv = rh.make_kv(model, tcfg, device)
    launches = None
    for i in range(max(1, warmup)):
        r = rh.generate(model, kv, ids, emb, mask, max_new_tokens=GEN_TOKENS - 1, count_launches=count_launches and i == 0)
        launches = r["launches_per_cycle"] if r["launches_per_cycle"] is not None else launches
    gi = torch.device(device).index or 0
    clocks = ClockSampler(gi)
    clocks.start()
    tokens, dec_s, pre_s, acc, out = 0, 0.0, 0.0, [], None
    for _ in range(steps):
        r = rh.generate(model, kv, ids, emb, mask, max_new_tokens=GEN_TOKENS - 1)
        tokens += r["new_token"]; dec_s += r["decode_s"]; pre_s += r["prefill_s"]; acc += r["accept"]
        out = r["output_ids"]
    clk = clocks.stop()
    res = {"value": tokens / dec_s, "unit": "tokens/s", "kind": "reference",
           "what": "unmodified vispec.model modules (baseline/_ref) on cuda, eager PyTorch " + torch.__version__,
           "steps": steps, "warmup": max(1, warmup), "tokens": tokens, "decode_s": dec_s, "ms_per_cycle": dec_s / max(1, len(acc)) * 1e3,
           "prefill_ms": pre_s / steps * 1e3, "accept_len_mean": sum(acc) / max(1, len(acc)),
           "kernel_launches_per_cycle": launches, "clocks": clk, "output_ids": out[0].tolist()}
    del model, kv, tw, dw
    torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------------------------------------------------
# algorithmic bytes of one draft+verify cycle (DESIGN.md "Roofline")
# ------------------------------------------------------------------------------------------------------------
def cycle_algorithmic_bytes(tcfg, dcfg, kv_len_target: float, kv_len_draft: float):
    H, I, V, nh, nkv, Lyr = tcfg.hidden, tcfg.inter, tcfg.vocab, tcfg.heads, tcfg.kv_heads, tcfg.layers
    w_target = Lyr * ((nh + 2 * nkv) * 128 * H + H * nh * 128 + 3 * H * I) * 2 + V * H * 2
    Hd, Id = dcfg.hidden, dcfg.inter
    w_draft_fwd = (2 * Hd * 2 * Hd + 4 * Hd * Hd + 3 * Hd * Id + V * Hd) * 2
    kv_target = Lyr * 2 * kv_len_target * nkv * 128 * 2
    kv_draft_fwd = 2 * kv_len_draft * Hd * 2
    n_fwd = dcfg.depth + 1
    return {"weights_target": w_target, "weights_draft": n_fwd * w_draft_fwd, "kv_target": kv_target,
            "kv_draft": n_fwd * kv_draft_fwd,
            "total": w_target + n_fwd * w_draft_fwd + kv_target + n_fwd * kv_draft_fwd}


def phase_trace(model, ids, emb, mask, cycles=4):
    """Per-phase durations of one cycle of the persistent kernel (us, CTA 0's %globaltimer stamps)."""
    from vispec_b200 import _lib as L
    from vispec_b200.model import utils as vu

    model.spec_layer.reset_kv()
    model.current_length_data.zero_()
    vu.reset_tree_mode(model)
    vu.initialize_tree(ids, model, model.past_key_values, None, emb, None, image_mask=mask)
    torch.cuda.synchronize()
    prog = model.cycle_program(trace=True)
    for _ in range(cycles):
        prog.run()
    torch.cuda.synchronize()
    us = prog.read_trace_us()
    types = prog.phase_types()
    by = {}
    for t, d in zip(types, us):
        k = L.PHASE_NAMES[t]
        n, tot = by.get(k, (0, 0.0))
        by[k] = (n + 1, tot + d)
    return types, us, by

# ------------------------------------------------------------------------------------------------------------
def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation on the host cores (the tier's reference arm).
    `--impl reference-gpu`: the reference's own eager PyTorch path on the B200 (the north star's denominator)."""
    rank, world, local = dist_setup(args.gpus)
    if rank != 0:
        return
    t_all0 = time.perf_counter()
    cfg = {"workload": WORKLOAD, "tree": "total_token=30 depth=3 top_k=8 num_q=2", "weights": "random-init"}
    if args.impl == "reference-gpu":
        device = f"cuda:{local}"
        torch.cuda.set_device(local)
        r = gpu_reference_run(device, args.steps, max(args.warmup, 1), seed=0, layers=args.layers)
        if "unavailable" in r:
            print(json.dumps({"impl": "reference-gpu", "unavailable": r["unavailable"]}), flush=True)
            return
        r.pop("output_ids", None)
        line = {"metric": "decode_tokens_per_sec", "value": r["value"], "unit": "tokens/s", "n_gpus": 1, "steps": args.steps,
                "warmup": r["warmup"], "ms_per_step": r["decode_s"] / args.steps * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "impl": "reference-gpu",
                "config": dict(cfg, path="reference eager PyTorch on 1xB200, decode region of spec_model_ours.py:480-482"),
                "accept_len_mean": r["accept_len_mean"], "gpu_eager_baseline": r, "clocks": r["clocks"],
                "e2e": {"value": r["tokens"] / (r["decode_s"] + r["prefill_ms"] * 1e-3 * args.steps), "unit": "tokens/s",
                        "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "wall_s": time.perf_counter() - t_all0}
        print(json.dumps(line), flush=True)
        return
    threads = os.cpu_count() or 1
    per_step = max(1, args.ref_cycles)
    warm = min(max(args.warmup, 0), 2) * per_step            # warm-up CYCLES actually run (bounded: ~0.3 s each)
    tokens, dt, acc, kind, what = cpu_reference_sample(per_step * args.steps, warm=warm, threads=threads)
    tps = tokens / dt
    line = {"metric": "decode_tokens_per_sec", "value": tps, "unit": "tokens/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": warm // per_step, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "impl": "reference",
            "config": dict(cfg, region="decode only (synthetic KV state, CPU prompt pass skipped)",
                           sample=f"{per_step} draft+verify cycle(s) per step"),
            "accept_len_mean": sum(acc) / max(1, len(acc)),
            "cpu_baseline": {"value": tps, "unit": "tokens/s", "cores": threads, "kind": kind,
                             "sample": f"{per_step} draft+verify cycle(s) per step after {warm} warm-up cycle(s) at S=704+ with "
                                       f"synthetic KV state (decode only; CPU prompt pass skipped); {what}; torch "
                                       f"{torch.__version__} CPU bf16"},
            "e2e": {"value": tps, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t_all0}
    print(json.dumps(line), flush=True)


def run_ours(args):
    rank, world, local = dist_setup(args.gpus)
    torch.cuda.set_device(local)
    device = f"cuda:{local}"
    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(device))
    from vispec_b200 import _lib as L

    lib = L.load()
    model, ids, emb, mask = build_llava7b(device=device, seed=rank, layers=args.layers)
    eng = model.base_model_engine
    gen_kw = dict(temperature=0.0, log=True, return_acceptance_len=True, max_new_tokens=GEN_TOKENS - 1)

    last_out = {}

    def decode_step():
        out, new_token, idx, acc = model.specgenerate(ids, inputs_embeds=emb, image_mask=mask, **gen_kw)
        last_out["ids"] = out
        return new_token, acc, model.last_decode_ms

    # pinned host copies for the e2e leg
    ids_h, emb_h, mask_h = ids.cpu().pin_memory(), emb.cpu().pin_memory(), mask.cpu().pin_memory()
    h2d_bytes = ids_h.numel() * 8 + emb_h.numel() * 2 + mask_h.numel()

    def e2e_step():
        i_d = ids_h.to(device, non_blocking=True)
        e_d = emb_h.to(device, non_blocking=True)
        m_d = mask_h.to(device, non_blocking=True)
        out, new_token, idx, acc = model.specgenerate(i_d, inputs_embeds=e_d, image_mask=m_d, **gen_kw)
        out_h = out.cpu()
        return new_token, out_h.numel() * 8

    for _ in range(max(args.warmup, 3)):
        decode_step()
    if use_dist:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = ClockSampler(local)
    clocks.start()
    lib.vs_launch_count(1)
    tokens, dec_ms, accs = 0, 0.0, []
    for _ in range(args.steps):
        n, acc, ms = decode_step()
        tokens += n
        dec_ms += ms
        accs += acc
    torch.cuda.synchronize()
    launches = lib.vs_launch_count(1)
    # e2e leg (host inputs, prefill inside)
    e2e_tokens, d2h = 0, 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        n, b = e2e_step()
        e2e_tokens += n
        d2h = b
    ev1.record()
    torch.cuda.synchronize()
    e2e_s = max(time.perf_counter() - t0, ev0.elapsed_time(ev1) / 1e3)
    clk = clocks.stop()
    if use_dist:
        dist.barrier()
    t = torch.tensor([dec_ms, e2e_s], dtype=torch.float64, device=device)
    cnt = torch.tensor([tokens, e2e_tokens, launches], dtype=torch.float64, device=device)
    if use_dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    dec_ms_max, e2e_s_max = t.tolist()
    tokens_all, e2e_tokens_all, launches_all = cnt.tolist()

    # roofline leg (rank 0): the dominant kernel is the persistent program kernel, one launch per cycle
    roof, roof_attn, phases = None, None, None
    if rank == 0:
        from vispec_b200.engine import use_programs
        peak, peak_src = hbm_peak()
        tcfg, dcfg = eng.cfg, model.spec_layer.config
        n_cycles = len(accs)
        prompt_len = ids.shape[1]
        mean_accept = sum(accs) / max(1, len(accs))
        kv_t = prompt_len + (GEN_TOKENS / 2.0) + dcfg.total_tokens            # mean target KV length seen by verify
        kv_d = (prompt_len - N_IMG + dcfg.num_q - 1) + (GEN_TOKENS / 2.0) + 8 * (dcfg.depth + 1) / 2.0
        ab = cycle_algorithmic_bytes(tcfg, dcfg, kv_t, kv_d)
        if use_programs() and n_cycles:
            avg_ms = dec_ms / n_cycles                                          # CUDA events around the timed region
            ach = ab["total"] / (avg_ms * 1e-3) / 1e9
            roof = {"bound": "hbm", "kernel": "program_kernel (persistent draft+verify cycle, one launch per cycle)",
                    "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": NCU_TRAFFIC_BYTES, "traffic_source": NCU_TRAFFIC_SOURCE, "peak_source": peak_src,
                    "launches_timed": n_cycles, "avg_launch_us": avg_ms * 1e3, "algorithmic_bytes_per_launch": ab["total"],
                    "algorithmic_bytes": ab, "mean_accept_len": mean_accept}
            types, us, by = phase_trace(model, ids, emb, mask)
            tot = sum(us)
            phases = {k: {"n": n, "us": round(t, 1), "share": round(t / tot, 4)} for k, (n, t) in sorted(by.items(), key=lambda kv: -kv[1][1])}
            n_attn = by.get("attn_scores", (0, 0.0))[0] + by.get("attn_fused", (0, 0.0))[0]
            if n_attn:
                # target layers only: the first `layers` attention phase pairs of the cycle
                t_attn = 0.0
                seen = 0
                n_target = tcfg.layers * (1 if by.get("attn_fused") else 2)
                for t, d in zip(types, us):
                    if t in (L.PH_ATTN_S, L.PH_ATTN_PV, L.PH_ATTN_FUSED) and seen < n_target:
                        t_attn += d
                        seen += 1
                kv_len = prompt_len + dcfg.total_tokens
                per_layer = 2 * kv_len * tcfg.kv_heads * 128 * 2
                ach_a = per_layer * tcfg.layers / (t_attn * 1e-6) / 1e9
                roof_attn = {"bound": "hbm", "kernel": "attention phase(s) of program_kernel (scores + softmax + PV, incl. the closing grid barrier)",
                             "achieved": ach_a, "peak": peak, "unit": "GB/s", "frac": ach_a / peak, "kv_len": kv_len,
                             "kv_bytes_per_layer": per_layer, "avg_layer_us": t_attn / tcfg.layers}
    # accept-length leg (rank 0, N=1): the same kernel on the bigram fixture of SURVEY.md 7.2 at the same shapes, where
    # the draft is (mostly) right -- tokens/s scales with accept length + 1 at an unchanged cycle time
    fixture = None
    if rank == 0 and world == 1 and not args.no_fixture:
        del model
        torch.cuda.empty_cache()
        m2, ids2, emb2, mask2 = build_llava7b(device=device, seed=rank + 3, layers=args.layers, fixture="bigram")
        for _ in range(2):
            m2.specgenerate(ids2, inputs_embeds=emb2, image_mask=mask2, **gen_kw)
        tok2, ms2, acc2 = 0, 0.0, []
        for _ in range(max(1, args.steps // 2)):
            _, n2, _, a2 = m2.specgenerate(ids2, inputs_embeds=emb2, image_mask=mask2, **gen_kw)
            tok2 += n2; ms2 += m2.last_decode_ms; acc2 += a2
        fixture = {"weights": "bigram fixture (target o_proj/down_proj = 0, draft = target embedding + identity fc)",
                   "value": tok2 / (ms2 * 1e-3), "unit": "tokens/s", "accept_len_mean": sum(acc2) / max(1, len(acc2)),
                   "ms_per_cycle": ms2 / max(1, len(acc2))}
        del m2
        torch.cuda.empty_cache()
    # second model family (rank 0, N=1, informational): the same kernel on Qwen2.5-VL-7B shapes (BASELINE config 3)
    qwen = None
    if rank == 0 and world == 1 and not args.no_fixture:
        m3, ids3, emb3, mask3, pos3, tc3, dc3 = build_qwen7b(device=device, seed=11)
        kw3 = dict(gen_kw, position_ids=pos3)
        for _ in range(2):
            m3.specgenerate(ids3, inputs_embeds=emb3, image_mask=mask3, **kw3)
        tok3, ms3, acc3 = 0, 0.0, []
        for _ in range(max(1, args.steps // 2)):
            _, n3, _, a3 = m3.specgenerate(ids3, inputs_embeds=emb3, image_mask=mask3, **kw3)
            tok3 += n3; ms3 += m3.last_decode_ms; acc3 += a3
        ms_cycle3 = ms3 / max(1, len(acc3))
        by3 = cycle_algorithmic_bytes(tc3, dc3, ids3.shape[1] + GEN_TOKENS / 2 + 31, 130 + GEN_TOKENS / 2)
        qwen = {"workload": "Qwen2.5-VL-7B shapes (28q/4kv GQA, qkv bias, V=152064) + ViSpec head, T=0, 576 img + 128 text "
                            "tokens, 256-token gen", "value": tok3 / (ms3 * 1e-3), "unit": "tokens/s",
                "accept_len_mean": sum(acc3) / max(1, len(acc3)), "ms_per_cycle": ms_cycle3,
                "algorithmic_bytes_per_cycle": by3["total"], "hbm_frac": by3["total"] / (ms_cycle3 * 1e-3) / 1e9 / hbm_peak()[0]}
        # BASELINE config 3 proper: T = 1.0 (SpecInfer-style acceptance + sampling on the device, Philox uniforms)
        kw3s = dict(kw3, temperature=1.0)
        torch.manual_seed(0)
        for _ in range(2):
            m3.specgenerate(ids3, inputs_embeds=emb3, image_mask=mask3, **kw3s)
        tok4, ms4, acc4 = 0, 0.0, []
        for _ in range(max(1, args.steps // 2)):
            _, n4, _, a4 = m3.specgenerate(ids3, inputs_embeds=emb3, image_mask=mask3, **kw3s)
            tok4 += n4; ms4 += m3.last_decode_ms; acc4 += a4
        qwen["sampling_t1"] = {"temperature": 1.0, "value": tok4 / (ms4 * 1e-3), "unit": "tokens/s",
                               "accept_len_mean": sum(acc4) / max(1, len(acc4)), "ms_per_cycle": ms4 / max(1, len(acc4))}
        del m3
        torch.cuda.empty_cache()
    # the verify attention on its own at longer contexts (rank 0, N=1, informational): 32 attention phases over distinct
    # K/V slabs per launch (tools/bench_attention.py); beyond 1152 keys the single-phase attention streams its chunks
    attn_long = None
    if rank == 0 and world == 1 and not args.no_fixture and use_programs():
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools"))
        from bench_attention import measure as attn_measure
        attn_long = [attn_measure(kv, True, peak=hbm_peak()[0], dev=device) for kv in (2048, 8192)]
        torch.cuda.empty_cache()
    if rank != 0:
        if use_dist:
            dist.destroy_process_group()
        return
    value = tokens_all / (dec_ms_max * 1e-3)
    e2e_v = e2e_tokens_all / e2e_s_max
    # the north star's denominator (rank 0, N=1): the reference's own eager PyTorch path on this B200
    gpu_ref = None
    if world == 1 and not args.no_gpu_baseline:
        gpu_ref = gpu_reference_run(device, steps=max(1, min(2, args.steps)), warmup=1, seed=rank, layers=args.layers)
        if "unavailable" not in gpu_ref:
            ref_ids = gpu_ref.pop("output_ids")
            mine = last_out["ids"][0].tolist()
            n_cmp = min(len(mine), len(ref_ids))
            same = 0
            while same < n_cmp and mine[same] == ref_ids[same]:
                same += 1
            gpu_ref["tokens_identical_to_ours"] = {"compared": n_cmp - ids.shape[1], "common_prefix": same - ids.shape[1]}
            gpu_ref["speedup_decode"] = value / gpu_ref["value"]
            gpu_ref["target_1p5x_met"] = bool(value >= 1.5 * gpu_ref["value"])
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        ctok, cdt, cacc, ckind, cwhat = cpu_reference_sample(args.ref_cycles, warm=1, threads=threads)
        cpu = {"value": ctok / cdt, "unit": "tokens/s", "cores": threads, "kind": ckind,
               "sample": f"{args.ref_cycles} draft+verify cycle(s) after 1 warm-up cycle at S=704+ with synthetic KV state "
                         f"(decode only); {cwhat}; CPU bf16"}
    line = {"metric": "decode_tokens_per_sec", "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dec_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "tree": "total_token=30 depth=3 top_k=8 num_q=2", "weights": "random-init",
                       "layers": args.layers, "l2": "inputs larger than L2 (16.4 GB of weights streamed per cycle)",
                       "path": "persistent program kernel" if os.environ.get("VS_MULTI_KERNEL", "0") != "1" else "one launch per op",
                       "parallelism": f"replicas x{world}"},
            "accept_len_mean": sum(accs) / max(1, len(accs)), "cycles_per_step": len(accs) / args.steps,
            "e2e": {"value": e2e_v, "unit": "tokens/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches_all), "clocks": clk, "roofline": roof, "roofline_attention": roof_attn,
            "phases_us_per_cycle": phases, "accept_fixture": fixture, "qwen2_5_vl_7b": qwen,
            "attention_long_context": attn_long, "gpu_eager_baseline": gpu_ref, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    if use_dist:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------------
# BASELINE config 5: eight independent batch-1 streams with mixed prompt lengths (576 image + 64..512 text tokens)
# ------------------------------------------------------------------------------------------------------------
MIXED_TEXT = (64, 128, 256, 512, 96, 192, 384, 448)       # text tokens of stream i (drawn once from {64 ... 512})
MIXED_NOISE = (0.0, 0.6, 0.3, 0.9, 0.45, 0.15, 0.75, 1.2)  # draft quality of stream i -> accept length varies


def run_mixed(args):
    """`--workload mixed`: the reference's multi-GPU scheme (prompt sharding over replicas, no data-path collective,
    gen_spec_answer_coco_caption.py:53-83) on BASELINE config 5.  Stream i = 576 image rows + MIXED_TEXT[i] text tokens
    on the LLaVA-7B bigram fixture with draft noise MIXED_NOISE[i]; rank r of N runs streams r, r + N, ...; 256 tokens
    per stream at T = 0.  value = all generated tokens / the slowest rank's decode time."""
    from collections import Counter

    from vispec_b200 import _lib as L
    from vispec_b200 import replicas

    rank, world, local = dist_setup(args.gpus)
    torch.cuda.set_device(local)
    device = f"cuda:{local}"
    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(device))
    lib = L.load()
    mine = list(replicas.shard_prompts(len(MIXED_TEXT), rank, world))
    gen_kw = dict(temperature=0.0, log=True, return_acceptance_len=True, max_new_tokens=GEN_TOKENS - 1)
    records, launches = [], 0
    clocks = ClockSampler(local)
    for si in mine:
        model, ids, emb, mask = build_llava7b(device=device, seed=100 + si, layers=args.layers, fixture="bigram",
                                              fc_noise=MIXED_NOISE[si], n_post=MIXED_TEXT[si] - 5)
        for _ in range(max(args.warmup, 1)):
            model.specgenerate(ids, inputs_embeds=emb, image_mask=mask, **gen_kw)
        torch.cuda.synchronize()
        if si == mine[0]:
            clocks.start()
        lib.vs_launch_count(1)
        tok, ms, acc, wall = 0, 0.0, [], 0.0
        for _ in range(args.steps):
            t0 = time.perf_counter()
            _, n, _, a = model.specgenerate(ids, inputs_embeds=emb, image_mask=mask, **gen_kw)
            torch.cuda.synchronize()
            wall += time.perf_counter() - t0
            tok += n; ms += model.last_decode_ms; acc += a
        launches += lib.vs_launch_count(1)
        records.append({"question_id": si, "text_tokens": MIXED_TEXT[si], "prompt_tokens": int(ids.shape[1]),
                        "fc_noise": MIXED_NOISE[si], "new_tokens": [tok], "wall_time": [wall], "decode_ms": ms,
                        "acceptance_length": acc, "tokens_per_s": tok / (ms * 1e-3), "rank": rank})
        del model
        torch.cuda.empty_cache()
    clk = clocks.stop()
    t = torch.tensor([sum(r["decode_ms"] for r in records)], dtype=torch.float64, device=device)
    cnt = torch.tensor([sum(r["new_tokens"][0] for r in records), launches], dtype=torch.float64, device=device)
    if use_dist:
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    allrec = replicas.gather_results(records)
    if rank == 0:
        hist = Counter(a for r in allrec for a in r["acceptance_length"])
        agg = replicas.aggregate_throughput(allrec)
        wall_by_rank = {}
        for r in allrec:
            wall_by_rank[r["rank"]] = wall_by_rank.get(r["rank"], 0.0) + sum(r["wall_time"])
        wall_max = max(wall_by_rank.values())
        line = {"metric": "decode_tokens_per_sec", "value": cnt[0].item() / (t.item() * 1e-3), "unit": "tokens/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": t.item() / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": "BASELINE config 5: 8 independent batch-1 streams, LLaVA-1.6-7B shapes + ViSpec head, 576 "
                                       "image + 64..512 text tokens, T=0, 256-token gen, bigram fixture with per-stream draft noise",
                           "parallelism": f"replicas x{world} (prompt sharding, no data-path collective)",
                           "l2": "inputs larger than L2 (16.4 GB of weights streamed per cycle)"},
                "streams": [{k: r[k] for k in ("question_id", "rank", "text_tokens", "prompt_tokens", "fc_noise", "tokens_per_s")}
                            | {"accept_len_mean": sum(r["acceptance_length"]) / max(1, len(r["acceptance_length"]))}
                            for r in allrec],
                "accept_len_hist": {str(k): hist[k] for k in sorted(hist)}, "accept_len_mean": agg["mean_accept_len"],
                "mean_stream_tokens_per_s": agg["mean_tokens_per_s"], "gpu_launches": int(cnt[1].item()), "clocks": clk,
                "e2e": {"value": sum(r["new_tokens"][0] for r in allrec) / wall_max,
                        "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                        "note": "wall clock around specgenerate incl. the prompt pass; inputs resident"}}
        print(json.dumps(line), flush=True)
    if use_dist:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------------
# BASELINE config 4: LLaVA-1.6-13B shapes, tensor-parallel target prompt pass over the N GPUs, replicated decode
# ------------------------------------------------------------------------------------------------------------
def run_tp13b(args):
    """`--workload llava13b_tp`: 13B target (H=5120, 40 heads, I=13824, 40 layers) + ViSpec head.  The prompt pass is
    sharded tensor-parallel over the N ranks (vispec_b200/tp_prefill.py: column-parallel q|k|v / gate|up, row-parallel
    o / down with one all-reduce each, KV all-gathered); the draft head and the draft-and-verify decode are replicated
    (one persistent kernel per GPU, every rank produces the same tokens).  Reports the prompt pass at 704 and 2885 rows
    (TP vs the single-GPU native pass on rank 0) and decode tokens/s; value = ONE stream's decode tokens/s (the ranks
    decode in lockstep: replicas of the same stream, not additional throughput)."""
    from vispec_b200.engine import DraftConfig, TargetConfig
    from vispec_b200.model.spec_model_ours import SpecModel

    rank, world, local = dist_setup(args.gpus)
    torch.cuda.set_device(local)
    device = f"cuda:{local}"
    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(device))
    layers = args.layers if args.layers != 32 else 40
    tcfg = TargetConfig(hidden=5120, inter=13824, layers=layers, heads=40, kv_heads=40, vocab=32064)
    dcfg = DraftConfig(hidden=5120, inter=13824, heads=40, vocab=32064)
    tw = synth_target_weights(tcfg, 5, device)          # same seed on every rank: identical replicas
    dw = synth_draft_weights(dcfg, 6, device)
    model = SpecModel.from_state_dicts(tcfg, tw, dcfg, dw, device=device, kv_max=4096, eos_token_id=None)
    del tw, dw
    torch.cuda.empty_cache()
    eng = model.base_model_engine
    gen_kw = dict(temperature=0.0, log=True, return_acceptance_len=True, max_new_tokens=GEN_TOKENS - 1)
    out = {}
    prefill = []
    for n_post in (123, 2304):                              # 704 rows; 5 + 576 + 2304 = 2885 rows
        ids, emb, mask = synth_prompt(tcfg, eng.embed, 12, device, n_post=n_post)
        rows = ids.shape[1]
        res = {"rows": rows}
        for mode in (("native_1gpu", "tp") if use_dist else ("native_1gpu",)):
            eng._tp = None
            if mode == "tp":
                model.enable_tensor_parallel_prefill()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            times = []
            for it in range(2 + max(1, args.steps)):
                eng.set_len(0)
                if use_dist:
                    dist.barrier()
                torch.cuda.synchronize()
                ev0.record()
                lg, _ = eng.prefill(emb)
                ev1.record()
                torch.cuda.synchronize()
                if it >= 2:
                    times.append(ev0.elapsed_time(ev1))
            tm = torch.tensor([sum(times) / len(times)], dtype=torch.float64, device=device)
            if use_dist:
                dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            res[mode + "_ms"] = tm.item()
            res[mode + "_next_token"] = int(torch.argmax(lg))
            if mode == "tp" and getattr(eng, "_tp_timing", None):      # VS_TP_TIMING=1: ms per kind of step (rank 0)
                res["tp_breakdown_ms"] = eng._tp_timing
        if use_dist:
            res["speedup_tp_over_native"] = res["native_1gpu_ms"] / res["tp_ms"]
            res["next_token_equal"] = res["native_1gpu_next_token"] == res["tp_next_token"]
        prefill.append(res)
    # decode on the 704-row prompt (TP prompt pass when N > 1), all ranks in lockstep
    ids, emb, mask = synth_prompt(tcfg, eng.embed, 12, device)
    for _ in range(max(args.warmup, 1)):
        model.specgenerate(ids, inputs_embeds=emb, image_mask=mask, **gen_kw)
    if use_dist:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = ClockSampler(local)
    clocks.start()
    tok, ms, acc, last = 0, 0.0, [], None
    for _ in range(args.steps):
        o, n, _, a = model.specgenerate(ids, inputs_embeds=emb, image_mask=mask, **gen_kw)
        tok += n; ms += model.last_decode_ms; acc += a; last = o
    torch.cuda.synchronize()
    clk = clocks.stop()
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    same = torch.ones(1, dtype=torch.int64, device=device)
    if use_dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ref = last.clone()
        dist.broadcast(ref, src=0)
        same = torch.tensor([int(ref.shape == last.shape and bool(torch.equal(ref, last)))], dtype=torch.int64, device=device)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
    if rank == 0:
        by = cycle_algorithmic_bytes(tcfg, dcfg, ids.shape[1] + GEN_TOKENS / 2 + 30, 130 + GEN_TOKENS / 2)
        cyc_ms = t.item() / max(1, len(acc))
        line = {"metric": "decode_tokens_per_sec", "value": tok / (t.item() * 1e-3), "unit": "tokens/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": t.item() / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"BASELINE config 4: LLaVA-1.6-13B shapes ({layers} layers) + ViSpec head, T=0, 576 img + 128 "
                                       "text tokens, 256-token gen; tensor-parallel target prompt pass, replicated decode",
                           "parallelism": f"tp{world} prompt pass (NCCL all-reduce on the row-parallel boundaries) + replicated decode",
                           "l2": "inputs larger than L2 (26 GB of weights streamed per cycle)"},
                "accept_len_mean": sum(acc) / max(1, len(acc)), "ms_per_cycle": cyc_ms,
                "roofline": {"bound": "hbm", "achieved": by["total"] / (cyc_ms * 1e-3) / 1e9, "peak": hbm_peak()[0], "unit": "GB/s",
                             "frac": by["total"] / (cyc_ms * 1e-3) / 1e9 / hbm_peak()[0], "traffic": None,
                             "algorithmic_bytes_per_launch": by["total"]},
                "prompt_pass": prefill, "tokens_identical_on_all_ranks": bool(same.item()), "clocks": clk,
                "e2e": {"value": tok / (t.item() * 1e-3), "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                        "note": "decode region only (inputs resident); the prompt pass is reported separately"}}
        print(json.dumps(line), flush=True)
    if use_dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-gpu"])
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--ref-cycles", type=int, default=3, help="CPU cycles per bounded sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the reference-on-this-GPU (eager PyTorch) leg")
    ap.add_argument("--no-fixture", action="store_true", help="skip the accept-length (bigram fixture) leg")
    ap.add_argument("--workload", default="llava7b", choices=["llava7b", "mixed", "llava13b_tp"],
                    help="llava7b = BASELINE config 2 (the headline line); mixed = config 5 (8 streams, 64..512 text tokens, "
                         "prompt sharding over the GPUs); llava13b_tp = config 4 (13B shapes, tensor-parallel prompt pass)")
    args = ap.parse_args()
    if args.impl in ("reference", "reference-gpu"):
        run_reference(args)
    elif args.workload == "mixed":
        run_mixed(args)
    elif args.workload == "llava13b_tp":
        run_tp13b(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
