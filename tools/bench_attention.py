"""Decode (tree-verify) attention of the program kernel on its own: one launch = `layers` attention phases over distinct
K/V slabs (inputs larger than L2), 30 tree queries x 32 heads, timed with CUDA events.  Reports microseconds per layer and
the achieved fraction of the HBM roofline for the KV-read bytes (2 * S * n_kv * 128 * 2 B per layer).

    python tools/bench_attention.py --kv 734 2048 4096 8192 [--two-phase]
"""
import argparse
import ctypes as C
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vispec_b200 import _lib as L  # noqa: E402
from vispec_b200.engine import ScratchBuffers  # noqa: E402
from vispec_b200.program import Program  # noqa: E402


def tree_bits(T, dev):
    """a chain/branch mix: node i sees itself, the root and node i // 2's ancestors"""
    bits = []
    for i in range(T):
        b, j = 0, i
        while True:
            b |= 1 << j
            if j == 0:
                break
            j = (j - 1) // 2
        bits.append(b)
    return torch.tensor(bits, dtype=torch.int64, device=dev)


def measure(kv_len, fused, layers=32, nh=32, nkv=32, T=30, reps=20, dev="cuda", eager=True, peak=None, **opts):
    rows = (kv_len + 63) // 64 * 64
    sc = ScratchBuffers(nh * 128, 256, nh, 1000, 32, rows, dev)
    g = torch.Generator(device=dev).manual_seed(1)
    K = torch.randn(layers, nkv, rows, 128, generator=g, device=dev, dtype=torch.bfloat16)
    V = torch.randn(layers, nkv, rows, 128, generator=g, device=dev, dtype=torch.bfloat16)
    Q = torch.randn(nh, 32, 128, generator=g, device=dev, dtype=torch.bfloat16)
    O = torch.zeros(T, nh * 128, dtype=torch.bfloat16, device=dev)
    bits = tree_bits(T, dev)
    nctx = torch.tensor([kv_len - T], dtype=torch.int32, device=dev)
    args = []
    p = Program(dev, attn_fused=fused, **opts)
    for li in range(layers):
        a = L.AttnArgs()
        a.Q, a.q_head_stride, a.K, a.V, a.kv_head_stride = Q.data_ptr(), 32 * 128, K[li].data_ptr(), V[li].data_ptr(), rows * 128
        a.n_heads, a.n_kv_heads, a.Tq = nh, nkv, T
        a.n_ctx, a.n_ctx_host = nctx.data_ptr(), 0
        a.n_new, a.causal, a.bits, a.eager, a.scale = T, 0, bits.data_ptr(), int(eager), 1 / math.sqrt(128)
        a.scores, a.score_ld, a.stats, a.max_chunks = sc.scores.data_ptr(), sc.score_ld, sc.stats.data_ptr(), sc.max_chunks
        a.opart, a.pv_splits, a.counters, a.O, a.ldo = sc.opart.data_ptr(), sc.pv_splits, sc.attn_counters.data_ptr(), O.data_ptr(), nh * 128
        a.prefix_stable = 1          # the verify step's case: kv length and prefix rows are final when the launch starts
        args.append(a)
        p.add("attention", C.byref(a))
    p.commit()
    for _ in range(3):
        p.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        p.run()
    e1.record()
    torch.cuda.synchronize()
    us_layer = e0.elapsed_time(e1) * 1e3 / reps / layers
    kv_bytes = 2 * kv_len * nkv * 128 * 2
    gbs = kv_bytes / us_layer / 1e3
    out = {"kv_len": kv_len, "path": ("single-phase (TMA-staged K/V)" if fused else "two-phase") + (" " + str(opts) if opts else ""),
           "us_per_layer": round(us_layer, 2),
           "kv_bytes_per_layer": kv_bytes, "achieved_gbs": round(gbs, 1), "includes": "one grid barrier per phase"}
    if peak:
        out["frac_of_hbm_peak"] = round(gbs / peak, 4)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--kv", type=int, nargs="+", default=[734, 2048, 4096, 8192])
    ap.add_argument("--two-phase", action="store_true", help="also time the two-phase program")
    ap.add_argument("--heads", type=int, default=32)
    ap.add_argument("--kv-heads", type=int, default=32)
    ap.add_argument("--no-stage", action="store_true", help="also time the program with the prefix staging held back")
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    peak = None
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk)).get("hbm_gbs")
    for kv in a.kv:
        for fused in ([True, False] if a.two_phase else [True]):
            try:
                print(json.dumps(measure(kv, fused, nh=a.heads, nkv=a.kv_heads, peak=peak, layers=a.layers, reps=a.reps)), flush=True)
            except RuntimeError as e:        # the two-phase program needs a -DVS_TWO_PHASE_ATTN=1 build (VS_LIB_PATH)
                print(json.dumps({"kv_len": kv, "path": "two-phase", "skipped": str(e)[-120:]}), flush=True)
        if a.no_stage:
            print(json.dumps(measure(kv, True, nh=a.heads, nkv=a.kv_heads, peak=peak, layers=a.layers, reps=a.reps, attn_prefetch=0)), flush=True)
