"""Slab weight layout (csrc/slab.h): the plan the library computes and an independent restatement of pack -> stage copy
-> swizzled read -> accumulator lane -> output feature must reproduce x @ W^T exactly, for plain and paired epilogues,
packed blocks that start inside a swizzle atom, several passes, fewer rows than CTAs and ragged last stages.  CPU only:
the library call is host arithmetic (no GPU)."""
import numpy as np
import pytest

import slab_model as sm
from vispec_b200 import _lib as L

CASES = [  # N, K, epi, grid
    (4096, 256, L.EPI_RESID, 148),          # 28 / 27 rows per CTA: blocks of 3584 bytes, kc = 4
    (1000, 192, L.EPI_STORE, 148),          # 7 / 6 rows
    (100, 64, L.EPI_STORE, 148),            # fewer rows than CTAs
    (32064, 128, L.EPI_STORE, 148),         # two accumulators of 109 rows
    (12288, 128, L.EPI_HEADS, 148),         # 42 / 41 pairs: q | k | v heads
    (22016, 64, L.EPI_SWIGLU, 148),         # 75 / 74 gate | up pairs
    (152064, 64, L.EPI_STORE, 148),         # several passes
    (2560, 320, L.EPI_STORE, 16),           # small grid: 160 rows per CTA -> two accumulators, ragged last stage
]


@pytest.mark.parametrize("N,K,epi,grid", CASES)
def test_slab_plan_pack_and_read_model_reproduce_the_gemm(N, K, epi, grid):
    lib = L.load()
    p = sm.plan(lib, N, K, epi, grid)
    assert p is not None
    assert p["n_tiles"] * p["blk_smem"] * p["kc"] <= p["x_off"] and p["x_off"] % 1024 == 0
    assert p["x_off"] + p["kc"] * 4096 <= 32768 and p["rows_alloc"] <= 112
    assert lib.vs_slab_bytes(N, K, epi, grid) == p["grid"] * p["passes"] * p["nk"] * p["n_tiles"] * p["blk_bytes"]
    rng = np.random.default_rng(N + K)
    W = (rng.standard_normal((N, K)).astype(np.float32).view(np.uint32) >> 16).astype(np.uint16)     # bf16 bit patterns
    X = rng.standard_normal((3, K)).astype(np.float32)
    slab = sm.pack(p, W)
    Y = sm.emulate_gemm(p, slab, X, N)
    Wf = (W.astype(np.uint32) << 16).view(np.float32)
    ref = X.astype(np.float64) @ Wf.astype(np.float64).T
    np.testing.assert_allclose(Y, ref, rtol=1e-4, atol=1e-3)     # fp32 accumulation in the model; a wrong row is O(10)
