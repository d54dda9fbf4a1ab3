"""Tensor-parallel target prefill (vispec_b200/tp_prefill.py): the sharding + collectives on CPU over `gloo`
(world size 2 and 4) against the oracle's single-device prefill; the NCCL run on GPUs is tests/tp_prefill_gpu_check.py
(launched by torchrun; see test_tp_prefill_two_gpus)."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_partition():
    from vispec_b200.tp_prefill import check_shardable, shard_bounds
    for n in (1, 7, 40, 172, 216):
        for world in (1, 2, 3, 4, 8):
            parts = [shard_bounds(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
    check_shardable(32, 32, 11008, 8)          # LLaVA-1.6-7B: 172 gate|up groups over 8 ranks (uneven) is fine
    check_shardable(40, 40, 13824, 8)          # LLaVA-1.6-13B
    check_shardable(28, 4, 18944, 4)           # Qwen2.5-VL-7B: at most 4 ranks (4 kv heads)
    with pytest.raises(ValueError):
        check_shardable(28, 4, 18944, 8)
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _pack_layers(cfg, w):
    """the engine's packed per-layer weights (engine.TargetEngine.__init__) from reference state-dict keys, on CPU"""
    from vispec_b200.engine import pack_gate_up
    out = []
    for li in range(cfg.layers):
        p = f"model.layers.{li}."
        t = dict(ln_in=w[p + "input_layernorm.weight"],
                 w_qkv=torch.cat([w[p + f"self_attn.{n}_proj.weight"] for n in "qkv"], dim=0).contiguous(),
                 b_qkv=None, w_o=w[p + "self_attn.o_proj.weight"], ln_post=w[p + "post_attention_layernorm.weight"],
                 w_gateup=pack_gate_up(w[p + "mlp.gate_proj.weight"], w[p + "mlp.up_proj.weight"]),
                 w_down=w[p + "mlp.down_proj.weight"])
        if (p + "self_attn.q_proj.bias") in w:
            t["b_qkv"] = torch.cat([w[p + f"self_attn.{n}_proj.bias"] for n in "qkv"], dim=0).contiguous()
        out.append(t)
    return out


def _worker(rank, world, port, q, variant="llama"):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    torch.set_num_threads(2)
    from oracle import synth
    from oracle import vispec_oracle as vo
    from vispec_b200 import tp_prefill
    cfg = synth.TINY_T
    if variant == "qwen":      # GQA (8 q heads on 4 kv heads), q/k/v biases, SDPA attention semantics; 1-D positions
        cfg = vo.TargetCfg(hidden=1024, inter=640, layers=2, heads=8, kv_heads=4, vocab=1000, max_pos=512, qkv_bias=True,
                           attn_impl="sdpa", rope_theta=1e6)
    w = synth.make_target_weights(cfg, 11)
    ids, emb, mask = synth.make_prompt(cfg, w, 5, 4, 24, 9)
    ref = vo.TargetOracle(cfg, w)
    ref_logits, ref_hidden = ref.forward(inputs_embeds=emb)
    Lq = emb.shape[1]
    kv = torch.zeros_like(ref.kv)
    pos = torch.arange(Lq)
    cos, sin = ref.cos[pos][None, None], ref.sin[pos][None, None]
    h = tp_prefill.prefill_layers(emb[0], _pack_layers(cfg, w), kv, heads=cfg.heads, kv_heads=cfg.kv_heads, inter=cfg.inter,
                                  eps=cfg.eps, cos=cos, sin=sin, past=0, sdpa=cfg.attn_impl == "sdpa", rank=rank, world=world)
    hidden = tp_prefill._rms(h, w["model.norm.weight"], cfg.eps)
    # second segment appended behind the first (past > 0): the mask and the KV rows of the other ranks must be in place
    emb2 = emb[:, :7] * 0.5
    ref_logits2, ref_hidden2 = ref.forward(inputs_embeds=emb2)
    pos2 = torch.arange(Lq, Lq + 7)
    h2 = tp_prefill.prefill_layers(emb2[0], _pack_layers(cfg, w), kv, heads=cfg.heads, kv_heads=cfg.kv_heads,
                                   inter=cfg.inter, eps=cfg.eps, cos=ref.cos[pos2][None, None], sin=ref.sin[pos2][None, None],
                                   past=Lq, sdpa=cfg.attn_impl == "sdpa", rank=rank, world=world)
    hidden2 = tp_prefill._rms(h2, w["model.norm.weight"], cfg.eps)
    gathered = [torch.empty_like(hidden) for _ in range(world)]
    dist.all_gather(gathered, hidden.contiguous())
    q.put(dict(rank=rank,
               hid_exact=float((hidden == ref_hidden[0]).float().mean()),
               hid_err=float((hidden.float() - ref_hidden[0].float()).abs().max()),
               hid2_err=float((hidden2.float() - ref_hidden2[0].float()).abs().max()),
               kv_err=float((kv[..., :Lq + 7, :].float() - ref.kv[..., :Lq + 7, :].float()).abs().max()),
               kv_exact=float((kv[..., :Lq + 7, :] == ref.kv[..., :Lq + 7, :]).float().mean()),
               same_on_all_ranks=all(torch.equal(g, hidden) for g in gathered),
               argmax_ok=int(torch.argmax(torch.nn.functional.linear(hidden[-1], w["lm_head.weight"]).float())) ==
               int(torch.argmax(ref_logits[0, -1]))))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,variant", [(2, "llama"), (2, "qwen"), (4, "qwen")])
def test_tp_prefill_matches_oracle_gloo(world, variant):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 23500 + (os.getpid() * 7 + world * 13 + len(variant)) % 3000
    ps = [ctx.Process(target=_worker, args=(r, world, port, q, variant)) for r in range(world)]
    [p.start() for p in ps]
    res = [q.get(timeout=240) for _ in ps]
    [p.join(timeout=60) for p in ps]
    for r in res:
        # bf16 results: identical up to the fp32 summation order of the row-parallel products (one bf16 ulp at |x| < 8)
        assert r["hid_err"] <= 2 ** -4 and r["hid2_err"] <= 2 ** -4 and r["kv_err"] <= 2 ** -4, r
        assert r["hid_exact"] > 0.5 and r["kv_exact"] > 0.8, r      # CPU bf16 GEMMs: 1-ulp differences compound over the layers
        assert r["same_on_all_ranks"] and r["argmax_ok"], r


@pytest.mark.gpu
def test_tp_prefill_two_gpus():
    """NCCL run of the same path through TargetEngine / SpecModel on 2 GPUs (skipped on a 1-GPU box)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(24500 + os.getpid() % 2000), os.path.join(ROOT, "tests", "tp_prefill_gpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "TP_PREFILL_OK" in out.stdout
