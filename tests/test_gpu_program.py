"""Parity of the persistent program kernel (GPU): every `vs_program_add_*` phase against the same plain fp32 PyTorch
statements (with the reference's bf16 rounding points) that tests/test_gpu_kernels.py uses for the per-op launches,
chains of dependent phases inside one launch, and the integer phases bit-exactly against the stand-alone kernels."""
import ctypes as C
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

import test_gpu_kernels as tk  # noqa: E402
from vispec_b200 import _lib as L  # noqa: E402
from vispec_b200.engine import ScratchBuffers, pack_gate_up, rope_tables  # noqa: E402
from vispec_b200.program import Program  # noqa: E402

DEV, BF = tk.DEV, tk.BF
randn, rbf, assert_bf16_close = tk.randn, tk.rbf, tk.assert_bf16_close


def linear_args(x, w, bias, T, N, K, epi=L.EPI_STORE, y=None, ldy=None, R=None, **kw):
    a = L.LinearArgs()
    a.X, a.ldx, a.W, a.ldw, a.bias = x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0), L.ptr(bias)
    a.T, a.N, a.K, a.splits, a.epi = T, N, K, 1, epi
    if y is not None:
        a.Y, a.ldy = y.data_ptr(), ldy if ldy is not None else y.stride(0)
    if R is not None:
        a.R, a.ldr = R.data_ptr(), R.stride(0)
    for k, v in kw.items():
        setattr(a, k, v)
    return a


def run_program(build, **opts):
    p = Program(DEV, **opts)
    build(p)
    p.commit()
    p.run()
    torch.cuda.synchronize()
    return p


@pytest.mark.parametrize("T,N,K", [(1, 256, 256), (5, 1000, 256), (30, 4096, 4096), (32, 4096, 11008), (8, 32064, 4096),
                                   (30, 4096, 8192), (3, 128, 64),
                                   (30, 5120, 13824), (30, 15360, 5120),      # LLaVA-1.6-13B: down_proj, q|k|v
                                   (30, 4608, 3584), (8, 152064, 3584)])      # Qwen2.5-VL-7B: q|k|v (28q + 2x4kv heads), LM head
@pytest.mark.parametrize("bias", [False, True])
def test_program_linear_store(T, N, K, bias):
    x, w = randn(T, K, seed=1), randn(N, K, std=K ** -0.5, seed=2)
    b = randn(N, std=0.1, seed=3) if bias else None
    y = torch.zeros(T, N, dtype=BF, device=DEV)
    a = linear_args(x, w, b, T, N, K, y=y)
    p = run_program(lambda p: p.add("linear", C.byref(a)))
    assert p.phase_types() == [L.PH_GEMM, L.PH_EPI]
    ref = x.float() @ w.float().T
    if bias:
        ref = ref + b.float()
    assert_bf16_close(y, ref.to(BF))
    # a second launch of the same program: the barrier counters were left at zero, the result is bit-identical
    y1 = y.clone()
    y.zero_()
    p.run()
    torch.cuda.synchronize()
    assert torch.equal(y, y1)


@pytest.mark.parametrize("pf_ahead", [0, 4, 64])
def test_program_prefetch_distance_does_not_change_results(pf_ahead):
    T, N, K = 30, 4096, 4096
    x, w = randn(T, K, seed=4), randn(N, K, std=K ** -0.5, seed=5)
    y = torch.zeros(T, N, dtype=BF, device=DEV)
    a = linear_args(x, w, None, T, N, K, y=y)
    run_program(lambda p: p.add("linear", C.byref(a)), pf_ahead=pf_ahead)
    assert_bf16_close(y, (x.float() @ w.float().T).to(BF))


def test_program_linear_residual_in_place():
    T, N, K = 30, 4096, 4096
    x, w, r = randn(T, K, seed=6), randn(N, K, std=K ** -0.5, seed=7), randn(T, N, seed=8)
    y = r.clone()
    a = linear_args(x, w, None, T, N, K, epi=L.EPI_RESID, y=y, R=y)
    run_program(lambda p: p.add("linear", C.byref(a)))
    assert_bf16_close(y, (r.float() + rbf(x.float() @ w.float().T)).to(BF))


def test_program_linear_swiglu():
    T, H, I = 30, 4096, 11008
    x = randn(T, H, seed=9)
    gate, up = randn(I, H, std=H ** -0.5, seed=10), randn(I, H, std=H ** -0.5, seed=11)
    w = pack_gate_up(gate, up)
    y = torch.zeros(T, I, dtype=BF, device=DEV)
    a = linear_args(x, w, None, T, 2 * I, H, epi=L.EPI_SWIGLU, y=y)
    run_program(lambda p: p.add("linear", C.byref(a)))
    g, u = rbf(x.float() @ gate.float().T), rbf(x.float() @ up.float().T)
    assert_bf16_close(y, (rbf(torch.nn.functional.silu(g)) * u).to(BF), frac_exact=0.97)


@pytest.mark.parametrize("nh,nkv,T", [(32, 32, 30), (4, 2, 7), (2, 2, 32)])
def test_program_linear_heads_rope_and_kv_append(nh, nkv, T):
    H, D, max_rows, max_pos = nh * 128, 128, 256, 512
    x = randn(T, H, seed=12)
    w = randn((nh + 2 * nkv) * D, H, std=H ** -0.5, seed=13)
    b = randn((nh + 2 * nkv) * D, std=0.1, seed=14)
    cos, sin = rope_tables(D, max_pos, 10000.0, BF, DEV)
    pos = torch.randint(0, max_pos, (T,), dtype=torch.int32, generator=torch.Generator().manual_seed(T + 1)).to(DEV)
    base = torch.tensor([17], dtype=torch.int32, device=DEV)
    Q = torch.zeros(nh, 32, D, dtype=BF, device=DEV)
    Kc = torch.zeros(nkv, max_rows, D, dtype=BF, device=DEV)
    Vc = torch.zeros(nkv, max_rows, D, dtype=BF, device=DEV)
    a = linear_args(x, w, b, T, (nh + 2 * nkv) * D, H, epi=L.EPI_HEADS, n_q_tiles=nh, n_k_tiles=nkv, rope=1,
                    cos=cos.data_ptr(), sin=sin.data_ptr(), pos_ids=pos.data_ptr(), Qout=Q.data_ptr(),
                    q_head_stride=32 * D, Kout=Kc.data_ptr(), Vout=Vc.data_ptr(), kv_head_stride=max_rows * D,
                    kv_row_base=base.data_ptr(), kv_row_off=3)
    run_program(lambda p: p.add("linear", C.byref(a)))
    qkv = rbf(x.float() @ w.float().T + b.float())
    q, k, v = qkv.split([nh * D, nkv * D, nkv * D], dim=-1)
    c, s = cos[pos.long()].float()[:, None], sin[pos.long()].float()[:, None]

    def rope(t, n):
        t = t.view(T, n, D)
        rot = torch.cat((-t[..., 64:], t[..., :64]), dim=-1)
        return (rbf(t * c) + rbf(rot * s)).to(BF)

    assert_bf16_close(Q[:, :T].transpose(0, 1), rope(q, nh), frac_exact=0.97)
    assert_bf16_close(Kc[:, 20:20 + T].transpose(0, 1), rope(k, nkv), frac_exact=0.97)
    assert_bf16_close(Vc[:, 20:20 + T].transpose(0, 1), v.view(T, nkv, D).to(BF))
    assert int(Kc[:, :20].abs().sum()) == 0 and int(Kc[:, 20 + T:].abs().sum()) == 0     # nothing else touched


def test_program_chain_linear_norm_linear():
    """Three dependent steps in ONE launch: the activation producer must wait for the phase that writes its tile."""
    T, H, I = 30, 1024, 2048
    x = randn(T, H, seed=30)
    w1, w2 = randn(H, H, std=H ** -0.5, seed=31), randn(I, H, std=H ** -0.5, seed=32)
    nw = (1 + 0.1 * torch.randn(H)).to(BF).to(DEV)
    h = torch.zeros(32, H, dtype=BF, device=DEV)
    xn = torch.zeros(32, H, dtype=BF, device=DEV)
    y = torch.zeros(T, I, dtype=BF, device=DEV)
    a1 = linear_args(x, w1, None, T, H, H, y=h)
    a2 = linear_args(xn, w2, None, T, I, H, y=y)

    def build(p):
        p.add("linear", C.byref(a1))
        p.add("rmsnorm", h.data_ptr(), H, nw.data_ptr(), xn.data_ptr(), H, T, H, 1e-5)
        p.add("linear", C.byref(a2))

    p = run_program(build)
    assert p.n_phases == 5
    h_ref = (x.float() @ w1.float().T).to(BF)
    assert_bf16_close(h[:T], h_ref)
    hf = h[:T].float()                       # continue from the kernel's own h: isolates each step's rounding
    xn_ref = nw * (hf * torch.rsqrt(hf.pow(2).mean(-1, keepdim=True) + 1e-5)).to(BF)
    assert_bf16_close(xn[:T], xn_ref)
    assert_bf16_close(y, (xn[:T].float() @ w2.float().T).to(BF))
    for _ in range(3):                       # repeated launches stay bit-identical (deterministic k-order reduction)
        y1 = y.clone()
        p.run()
        torch.cuda.synchronize()
        assert torch.equal(y, y1)


def test_program_rows_argmax_topk():
    V, k = 32064, 8
    tab = randn(1000, 256, seed=16)
    idx = torch.tensor([5, 999, 0, 5], dtype=torch.int32, device=DEV)
    dst = torch.zeros(4, 512, dtype=BF, device=DEV)
    x = randn(30, V, std=4.0, seed=17)
    x[3, 100] = x[3, 20000] = 50.0
    amax = torch.zeros(30, dtype=torch.int32, device=DEV)
    tv = torch.zeros(8 * k, dtype=torch.float32, device=DEV)
    ti = torch.zeros(8 * k, dtype=torch.int32, device=DEV)
    sel = torch.tensor([5], dtype=torch.int32, device=DEV)
    tv1 = torch.zeros(k, dtype=torch.float32, device=DEV)
    ti1 = torch.zeros(k, dtype=torch.int32, device=DEV)

    def build(p):
        p.add("embed_rows", tab.data_ptr(), 256, idx.data_ptr(), dst.data_ptr(), 512, 4, 256, 1000)
        p.add("broadcast_row", tab[7].data_ptr(), dst[:, 256:].data_ptr(), 512, 4, 256)
        p.add("argmax_rows", x.data_ptr(), V, 30, V, amax.data_ptr())
        p.add("logsoftmax_topk", x.data_ptr(), V, 8, V, k, None, tv.data_ptr(), ti.data_ptr())
        p.add("logsoftmax_topk", x.data_ptr(), V, 1, V, k, sel.data_ptr(), tv1.data_ptr(), ti1.data_ptr())

    run_program(build)
    assert torch.equal(dst[:, :256], tab[idx.long()]) and torch.equal(dst[:, 256:], tab[7].expand(4, 256))
    assert int(amax[3]) == 100 and torch.equal(amax.long(), torch.argmax(x.float(), dim=-1))
    # the stand-alone kernel is the bit-exact reference of the phase (same formulas, same tie rule)
    tv_k = torch.zeros_like(tv)
    ti_k = torch.zeros_like(ti)
    L.check(L.load().vs_logsoftmax_topk(x.data_ptr(), V, 8, V, k, None, tv_k.data_ptr(), ti_k.data_ptr(), tk.stream()))
    torch.cuda.synchronize()
    lp = torch.log_softmax(x[:8].float(), dim=-1).to(BF).float()
    rv, ri = tk.ref_topk_lowest_index(lp, k)
    torch.testing.assert_close(tv.view(8, k), rv, rtol=0, atol=2 ** -5)
    gaps = (rv[:, :-1] - rv[:, 1:]).abs().min(dim=1).values
    for r in range(8):
        if gaps[r] > 2 ** -4:
            assert torch.equal(ti.view(8, k)[r].long(), ri[r])
    assert torch.equal(ti, ti_k)
    assert torch.equal(tv1, tv.view(8, k)[5]) and torch.equal(ti1, ti.view(8, k)[5])


def program_attention(Q, K, V, nh, nkv, Tq, n_ctx, n_new, causal, bits, eager, n_ctx_dev=True, fused=False, stable=0, **opts):
    sc = ScratchBuffers(nh * 128, 256, nh, 1000, 32, K.shape[1], DEV)
    O = torch.zeros(Tq, nh * 128, dtype=BF, device=DEV)
    a = L.AttnArgs()
    a.Q, a.q_head_stride, a.K, a.V, a.kv_head_stride = Q.data_ptr(), 32 * 128, K.data_ptr(), V.data_ptr(), K.shape[1] * 128
    a.n_heads, a.n_kv_heads, a.Tq = nh, nkv, Tq
    nctx_t = torch.tensor([n_ctx], dtype=torch.int32, device=DEV)
    a.n_ctx, a.n_ctx_host = (nctx_t.data_ptr(), 0) if n_ctx_dev else (None, n_ctx)
    a.n_new, a.causal, a.bits, a.eager, a.scale = n_new, int(causal), L.ptr(bits), int(eager), 1 / math.sqrt(128)
    a.scores, a.score_ld, a.stats, a.max_chunks = sc.scores.data_ptr(), sc.score_ld, sc.stats.data_ptr(), sc.max_chunks
    a.opart, a.pv_splits, a.counters, a.O, a.ldo = sc.opart.data_ptr(), sc.pv_splits, sc.attn_counters.data_ptr(), O.data_ptr(), nh * 128
    a.prefix_stable = stable
    if fused:
        assert n_ctx + n_new <= sc.fused_kv_capacity
    try:
        p = run_program(lambda p: p.add("attention", C.byref(a)), attn_fused=fused, **opts)
    except RuntimeError as e:
        if not fused and "two-phase attention program is not compiled" in str(e):
            pytest.skip("two-phase attention program: build with -DVS_TWO_PHASE_ATTN=1 (tools/build_variant.py)")
        raise
    assert p.phase_types() == ([L.PH_ATTN_FUSED] if fused else [L.PH_ATTN_S, L.PH_ATTN_PV])
    assert int(sc.attn_counters.abs().sum()) == 0
    if fused:                                 # relaunch: the per-head arrival counters are monotonic across launches
        O1 = O.clone()
        O.zero_()
        p.run()
        torch.cuda.synchronize()
        assert torch.equal(O, O1)
    return O


@pytest.mark.parametrize("eager", [True, False])
@pytest.mark.parametrize("nh,nkv,n_ctx,fused", [(32, 32, 704, False), (4, 2, 61, False), (2, 2, 0, False), (8, 8, 3000, False),
                                                (32, 32, 704, True), (32, 32, 1100, True), (4, 2, 61, True), (2, 2, 0, True),
                                                (8, 8, 1900, True),
                                                # beyond 2 chunks per split the single-phase attention streams its chunks
                                                (32, 32, 1200, True), (32, 32, 3000, True), (32, 32, 8100, True),
                                                (4, 2, 2500, True), (28, 4, 3333, True), (8, 8, 4000, True), (40, 40, 2000, True)])
def test_program_attention_tree(eager, nh, nkv, n_ctx, fused):
    T, rows = 30, max(4096, (n_ctx + 30 + 63) // 64 * 64)
    Q = randn(nh, 32, 128, seed=20)
    K, V = randn(nkv, rows, 128, seed=21), randn(nkv, rows, 128, seed=22)
    bits = tk.tree_bits(T, 5)
    O = program_attention(Q, K, V, nh, nkv, T, n_ctx, T, False, bits, eager, fused=fused)
    ref = tk.attention_ref(Q[:, :T], K, V, n_ctx, T, False, bits.tolist(), eager, nh, nkv)
    torch.testing.assert_close(O.float(), ref.float(), rtol=2 ** -6, atol=2e-2)
    assert (O.float() == ref.float()).float().mean() > 0.9


@pytest.mark.parametrize("eager", [True, False])
@pytest.mark.parametrize("opts", [dict(kv_two_loads=1), dict(attn_prefetch=0), dict()])
@pytest.mark.parametrize("nh,nkv,n_ctx", [(32, 32, 734), (28, 4, 700), (4, 2, 61), (2, 2, 0), (32, 32, 2047), (8, 8, 5000), (4, 4, 31),
                                          (4, 4, 32), (4, 4, 33), (32, 32, 1122)])
def test_program_attention_tma_staged_variants(eager, opts, nh, nkv, n_ctx):
    """The single-phase attention stages K/V through tensor-map TMA granules (attn_ring.cuh): the stable prefix staged ahead
    of the phase (prefix_stable = 1), the same with the staging held back (attn_prefetch = 0) and the two-loads-per-granule
    form of the tensor map must all give the result of the in-phase loads, bit for bit."""
    T, rows = 30, max(4096, (n_ctx + 30 + 63) // 64 * 64)
    Q = randn(nh, 32, 128, seed=30)
    K, V = randn(nkv, rows, 128, seed=31), randn(nkv, rows, 128, seed=32)
    K[:, n_ctx + T:] = float("nan")           # rows past the kv length may hold anything
    V[:, n_ctx + T:] = float("nan")
    bits = tk.tree_bits(T, 5)
    base = program_attention(Q, K, V, nh, nkv, T, n_ctx, T, False, bits, eager, fused=True)
    O = program_attention(Q, K, V, nh, nkv, T, n_ctx, T, False, bits, eager, fused=True, stable=1, **opts)
    assert torch.equal(O, base)
    ref = tk.attention_ref(Q[:, :T], K[:, :n_ctx + T], V[:, :n_ctx + T], n_ctx, T, False, bits.tolist(), eager, nh, nkv)
    torch.testing.assert_close(O.float(), ref.float(), rtol=2 ** -6, atol=2e-2)
    assert (O.float() == ref.float()).float().mean() > 0.9


def test_program_attention_phases_back_to_back_with_staging():
    """Several attention phases over different K/V slabs in one launch, all flagged prefix_stable: the KV producer runs
    ahead of the workers across phase boundaries; every phase must give the result of a program of its own."""
    nh = nkv = 32
    T, n_ctx, rows, layers = 30, 900, 1024, 6
    Q = randn(nh, 32, 128, seed=40)
    Ks = [randn(nkv, rows, 128, seed=41 + i) for i in range(layers)]
    Vs = [randn(nkv, rows, 128, seed=61 + i) for i in range(layers)]
    bits = tk.tree_bits(T, 5)
    single = [program_attention(Q, Ks[i], Vs[i], nh, nkv, T, n_ctx, T, False, bits, True, fused=True) for i in range(layers)]
    sc = ScratchBuffers(nh * 128, 256, nh, 1000, 32, rows, DEV)
    Os = [torch.zeros(T, nh * 128, dtype=BF, device=DEV) for _ in range(layers)]
    nctx_t = torch.tensor([n_ctx], dtype=torch.int32, device=DEV)
    args = []
    for i in range(layers):
        a = L.AttnArgs()
        a.Q, a.q_head_stride, a.K, a.V, a.kv_head_stride = Q.data_ptr(), 32 * 128, Ks[i].data_ptr(), Vs[i].data_ptr(), rows * 128
        a.n_heads, a.n_kv_heads, a.Tq = nh, nkv, T
        a.n_ctx, a.n_ctx_host = nctx_t.data_ptr(), 0
        a.n_new, a.causal, a.bits, a.eager, a.scale = T, 0, L.ptr(bits), 1, 1 / math.sqrt(128)
        a.scores, a.score_ld, a.stats, a.max_chunks = sc.scores.data_ptr(), sc.score_ld, sc.stats.data_ptr(), sc.max_chunks
        a.opart, a.pv_splits, a.counters, a.O, a.ldo = sc.opart.data_ptr(), sc.pv_splits, sc.attn_counters.data_ptr(), Os[i].data_ptr(), nh * 128
        a.prefix_stable = 1
        args.append(a)

    def build(p):
        for a in args:
            p.add("attention", C.byref(a))
    p = run_program(build, attn_fused=True)
    for _ in range(3):
        p.run()
    torch.cuda.synchronize()
    for i in range(layers):
        assert torch.equal(Os[i], single[i]), f"phase {i}"


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("Tq,n_ctx", [(1, 100), (5, 129), (8, 40), (32, 0), (1, 3000), (8, 2100), (32, 3990)])
def test_program_attention_causal(Tq, n_ctx, fused):
    nh = nkv = 4
    Q = randn(nh, 32, 128, seed=23)
    K, V = randn(nkv, 4096, 128, seed=24), randn(nkv, 4096, 128, seed=25)
    O = program_attention(Q, K, V, nh, nkv, Tq, n_ctx, Tq, True, None, False, n_ctx_dev=(n_ctx > 0), fused=fused)
    ref = tk.attention_ref(Q[:, :Tq], K, V, n_ctx, Tq, True, None, False, nh, nkv)
    torch.testing.assert_close(O.float(), ref.float(), rtol=2 ** -6, atol=2e-2)


@pytest.mark.parametrize("sel", [[0], [0, 3, 9, 29], [0, 2, 3, 4, 28]])
def test_program_kv_splice(sel):
    n_slabs, n_kv, max_pos, prev = 6, 4, 256, 100
    slab = randn(n_slabs, 1, n_kv, max_pos, 128, seed=29)
    ref = slab.clone()
    idx = torch.tensor(sel, device=DEV) + prev
    ref[..., prev:prev + len(sel), :] = ref[..., idx, :]                     # utils.py:529-541
    st = torch.zeros(L.ST_SLOTS, dtype=torch.int32, device=DEV)
    st[L.ST_PREV_LEN], st[L.ST_ACCEPT] = prev, len(sel) - 1
    sel_dev = torch.tensor(sel + [sel[-1]] * (L.VS_MAX_ACCEPT - len(sel)), dtype=torch.int32, device=DEV)
    run_program(lambda p: p.add("kv_splice", slab.data_ptr(), n_slabs, n_kv, max_pos, st.data_ptr(), sel_dev.data_ptr()))
    assert torch.equal(slab, ref)


def test_program_rejects_more_than_32_rows():
    x, w = randn(40, 256, seed=1), randn(256, 256, seed=2)
    y = torch.zeros(40, 256, dtype=BF, device=DEV)
    a = linear_args(x, w, None, 40, 256, 256, y=y)
    p = Program(DEV)
    with pytest.raises(RuntimeError, match="32"):
        p.add("linear", C.byref(a))
    with pytest.raises(RuntimeError, match="empty"):
        p.commit()


@pytest.mark.parametrize("T,N,K", [(30, 4096, 4096), (8, 32064, 4096), (5, 1000, 256), (32, 4096, 11008)])
@pytest.mark.parametrize("pf_ahead", [0, 16])
def test_program_linear_tiled_weights_bit_identical(T, N, K, pf_ahead):
    """The tiled HBM layout (contiguous, pre-swizzled 16 KB blocks + L2 prefetch) must give the same bits as the
    row-major tensor-map path: it only changes how the bytes travel."""
    from vispec_b200.engine import pack_weight_tiles
    x, w = randn(T, K, seed=41), randn(N, K, std=K ** -0.5, seed=42)
    b = randn(N, std=0.1, seed=43)
    y0 = torch.zeros(T, N, dtype=BF, device=DEV)
    y1 = torch.zeros(T, N, dtype=BF, device=DEV)
    a0 = linear_args(x, w, b, T, N, K, y=y0)
    run_program(lambda p: p.add("linear", C.byref(a0)), pf_ahead=0)
    wt = pack_weight_tiles(w)
    assert wt.data_ptr() % 1024 == 0 and wt.numel() == ((N + 127) // 128) * 128 * K
    a1 = linear_args(x, w, b, T, N, K, y=y1)
    run_program(lambda p: p.register_tiled([(w, wt)]).add("linear", C.byref(a1)), pf_ahead=pf_ahead)
    assert torch.equal(y0, y1)
    assert_bf16_close(y1, (x.float() @ w.float().T + b.float()).to(BF))


@pytest.mark.parametrize("hidden,heads,inter,T", [(4096, 32, 11008, 30), (5120, 40, 13824, 30), (3584, 28, 18944, 30),
                                                  (512, 4, 1024, 5), (4096, 32, 11008, 1)])
def test_program_rmsnorm_folded_into_residual_epilogue(monkeypatch, hidden, heads, inter, T):
    """An RMSNorm behind a residual epilogue runs inside that epilogue (per-group arrival counters, values still in
    registers): one phase fewer per norm, and bit-identical hidden states / logits to the program with PH_NORM phases
    (same per-tile sums of squares, same summation order), launch after launch."""
    from bench import synth_target_weights
    from vispec_b200.engine import TargetConfig, TargetEngine
    cfg = TargetConfig(hidden=hidden, inter=inter, layers=2, heads=heads, kv_heads=heads, vocab=1024)
    tw = synth_target_weights(cfg, 3, DEV)
    outs, phases = [], []
    for fuse in ("0", "1"):
        monkeypatch.setenv("VS_FUSE_NORM", fuse)
        eng = TargetEngine(cfg, tw, device=DEV, kv_max=256)
        eng.prefill(0.5 * randn(40, hidden, seed=5))
        ids = ((torch.arange(T, device=DEV) * 37 + 11) % cfg.vocab).to(torch.int32)
        pos = (40 + torch.arange(T, device=DEV)).to(torch.int32)
        for rep in range(3):                   # relaunches: the arrival counters are monotonic across launches
            eng.verify(ids, pos, None, T, cur_len=40)
            torch.cuda.synchronize()
            outs.append((eng.hidden_out[:T].clone(), eng.logits[:T].clone()))
        prog = next(iter(eng._verify_progs.values()))
        phases.append(prog.phase_types())
    n_norm = [p.count(L.PH_NORM) for p in phases]
    assert n_norm == [2 * cfg.layers + 1, 1] and len(phases[0]) - len(phases[1]) == 2 * cfg.layers
    for h, lg in outs[1:]:
        assert torch.equal(h, outs[0][0]) and torch.equal(lg, outs[0][1])
    assert bool(torch.isfinite(outs[0][1].float()).all()) and float(outs[0][0].float().abs().mean()) > 0.1


# ---------------------------------------------------------------------------------------------------------------------
# slab weights (csrc/slab.h): whole output rows per CTA, full-K accumulation, epilogue from TMEM inside the GEMM phase
# ---------------------------------------------------------------------------------------------------------------------
def slab_program(build, pairs, **opts):
    """pairs: [(row-major weight, epi)] -> registered slab images; returns the executed program."""
    from vispec_b200.engine import pack_weight_slab
    regs = []
    for w, epi in pairs:
        s = pack_weight_slab(w, epi)
        assert s is not None and s.data_ptr() % 128 == 0
        regs.append((w, s, epi))
    return run_program(lambda p: build(p.register_tiled(regs)), **opts)


@pytest.mark.parametrize("T,N,K", [(30, 4096, 4096),            # o_proj: 28 rows per CTA, 4 k-chunks per stage
                                   (32, 4096, 11008),           # down_proj: 172 k-chunks = 43 stages of 4
                                   (8, 32064, 4096),            # LM head: 2 accumulators of 109 rows
                                   (30, 4096, 8192),            # draft fc / img_fc
                                   (1, 256, 256), (5, 1000, 256), (3, 128, 64),   # fewer rows than CTAs; single k-chunk
                                   (7, 4096, 4160),             # 65 k-chunks: a last stage with fewer chunks than kc
                                   (30, 5120, 13824), (8, 152064, 3584)])        # 13B down_proj; Qwen LM head: several passes
@pytest.mark.parametrize("bias", [False, True])
def test_program_slab_linear_store(T, N, K, bias):
    x, w = randn(T, K, seed=51), randn(N, K, std=K ** -0.5, seed=52)
    b = randn(N, std=0.1, seed=53) if bias else None
    y = torch.full((T, N), 7.0, dtype=BF, device=DEV)
    a = linear_args(x, w, b, T, N, K, y=y)
    p = slab_program(lambda p: p.add("linear", C.byref(a)), [(w, L.EPI_STORE)])
    assert p.phase_types() == [L.PH_GEMM2]
    ref = x.float() @ w.float().T
    if bias:
        ref = ref + b.float()
    assert_bf16_close(y, ref.to(BF))
    y1 = y.clone()
    y.zero_()
    p.run()                                   # relaunch: barrier / arrival counters were left consistent
    torch.cuda.synchronize()
    assert torch.equal(y, y1)


@pytest.mark.parametrize("x3d", [0, 1])
def test_program_slab_activation_load_variants_agree(x3d):
    """One 3-D tensor-map load per stage (box {64, 32, kc}) and kc 2-D loads must deliver the same activation blocks."""
    T, N, K = 30, 4096, 4096
    x, w = randn(T, K, seed=54), randn(N, K, std=K ** -0.5, seed=55)
    y = torch.zeros(T, N, dtype=BF, device=DEV)
    a = linear_args(x, w, None, T, N, K, y=y)

    def build(p):
        p.set_option("x3d", x3d)
        p.add("linear", C.byref(a))

    slab_program(build, [(w, L.EPI_STORE)])
    assert_bf16_close(y, (x.float() @ w.float().T).to(BF))


@pytest.mark.parametrize("T", [30, 1])
def test_program_slab_residual_and_folded_norm(T):
    H, K = 4096, 11008
    x, w, r = randn(T, K, seed=56), randn(H, K, std=K ** -0.5, seed=57), randn(32, H, seed=58)
    nw = (1 + 0.1 * torch.randn(H)).to(BF).to(DEV)
    h = r.clone()
    xn = torch.zeros(32, H, dtype=BF, device=DEV)
    a = linear_args(x, w, None, T, H, K, epi=L.EPI_RESID, y=h, R=h)

    def build(p):
        p.add("linear", C.byref(a))
        p.add("rmsnorm", h.data_ptr(), H, nw.data_ptr(), xn.data_ptr(), H, T, H, 1e-5)

    p = slab_program(build, [(w, L.EPI_RESID)])
    assert p.phase_types() == [L.PH_GEMM2]                     # the norm ran inside the GEMM phase
    h_ref = (r[:T].float() + rbf(x.float() @ w.float().T)).to(BF)
    assert_bf16_close(h[:T], h_ref, atol=4e-2)      # (one ulp of a residual value in [4, 8) is 0.031)
    hf = h[:T].float()
    xn_ref = nw * (hf * torch.rsqrt(hf.pow(2).mean(-1, keepdim=True) + 1e-5)).to(BF)
    assert_bf16_close(xn[:T], xn_ref, frac_exact=0.98)
    for _ in range(2):                                         # the arrival counter is monotonic across launches
        h.copy_(r)
        p.run()
        torch.cuda.synchronize()
        assert_bf16_close(xn[:T], xn_ref, frac_exact=0.98)


def test_program_slab_swiglu():
    T, H, I = 30, 4096, 11008
    x = randn(T, H, seed=59)
    gate, up = randn(I, H, std=H ** -0.5, seed=60), randn(I, H, std=H ** -0.5, seed=61)
    w = pack_gate_up(gate, up)
    y = torch.zeros(T, I, dtype=BF, device=DEV)
    a = linear_args(x, w, None, T, 2 * I, H, epi=L.EPI_SWIGLU, y=y)
    p = slab_program(lambda p: p.add("linear", C.byref(a)), [(w, L.EPI_SWIGLU)])
    assert p.phase_types() == [L.PH_GEMM2]
    g, u = rbf(x.float() @ gate.float().T), rbf(x.float() @ up.float().T)
    assert_bf16_close(y, (rbf(torch.nn.functional.silu(g)) * u).to(BF), frac_exact=0.97)


@pytest.mark.parametrize("nh,nkv,T", [(32, 32, 30), (4, 2, 7), (2, 2, 32), (28, 4, 30)])
def test_program_slab_heads_rope_and_kv_append(nh, nkv, T):
    H, D, max_rows, max_pos = nh * 128, 128, 256, 512
    x = randn(T, H, seed=62)
    w = randn((nh + 2 * nkv) * D, H, std=H ** -0.5, seed=63)
    b = randn((nh + 2 * nkv) * D, std=0.1, seed=64)
    cos, sin = rope_tables(D, max_pos, 10000.0, BF, DEV)
    pos = torch.randint(0, max_pos, (T,), dtype=torch.int32, generator=torch.Generator().manual_seed(T)).to(DEV)
    base = torch.tensor([17], dtype=torch.int32, device=DEV)
    Q = torch.zeros(nh, 32, D, dtype=BF, device=DEV)
    Kc = torch.zeros(nkv, max_rows, D, dtype=BF, device=DEV)
    Vc = torch.zeros(nkv, max_rows, D, dtype=BF, device=DEV)
    a = linear_args(x, w, b, T, (nh + 2 * nkv) * D, H, epi=L.EPI_HEADS, n_q_tiles=nh, n_k_tiles=nkv, rope=1,
                    cos=cos.data_ptr(), sin=sin.data_ptr(), pos_ids=pos.data_ptr(), Qout=Q.data_ptr(),
                    q_head_stride=32 * D, Kout=Kc.data_ptr(), Vout=Vc.data_ptr(), kv_head_stride=max_rows * D,
                    kv_row_base=base.data_ptr(), kv_row_off=3)
    p = slab_program(lambda p: p.add("linear", C.byref(a)), [(w, L.EPI_HEADS)])
    assert p.phase_types() == [L.PH_GEMM2]
    qkv = rbf(x.float() @ w.float().T + b.float())
    q, k, v = qkv.split([nh * D, nkv * D, nkv * D], dim=-1)
    c, s = cos[pos.long()].float()[:, None], sin[pos.long()].float()[:, None]

    def rope(t, n):
        t = t.view(T, n, D)
        rot = torch.cat((-t[..., 64:], t[..., :64]), dim=-1)
        return (rbf(t * c) + rbf(rot * s)).to(BF)

    # (a rotated value is the sum of two rounded products: one input ulp can move a near-cancelling output by > 1e-2)
    assert_bf16_close(Q[:, :T].transpose(0, 1), rope(q, nh), frac_exact=0.97, atol=4e-2)
    assert_bf16_close(Kc[:, 20:20 + T].transpose(0, 1), rope(k, nkv), frac_exact=0.97, atol=4e-2)
    assert_bf16_close(Vc[:, 20:20 + T].transpose(0, 1), v.view(T, nkv, D).to(BF))
    assert int(Kc[:, :20].abs().sum()) == 0 and int(Kc[:, 20 + T:].abs().sum()) == 0     # nothing else touched


def test_program_slab_and_stream_k_phases_share_one_launch():
    """Slab GEMM -> stand-alone norm -> stream-K GEMM (unregistered weights) -> slab GEMM in ONE launch: the role warps
    walk both kinds of GEMM phase over the same ring and the same accumulator hand-shake."""
    T, H, I = 30, 1024, 2048
    x = randn(T, H, seed=65)
    w1, w2, w3 = randn(H, H, std=H ** -0.5, seed=66), randn(I, H, std=H ** -0.5, seed=67), randn(H, I, std=I ** -0.5, seed=68)
    nw = (1 + 0.1 * torch.randn(H)).to(BF).to(DEV)
    h = torch.zeros(32, H, dtype=BF, device=DEV)
    xn = torch.zeros(32, H, dtype=BF, device=DEV)
    y = torch.zeros(32, I, dtype=BF, device=DEV)
    z = torch.zeros(T, H, dtype=BF, device=DEV)
    a1 = linear_args(x, w1, None, T, H, H, y=h)
    a2 = linear_args(xn, w2, None, T, I, H, y=y)
    a3 = linear_args(y, w3, None, T, H, I, y=z)

    def build(p):
        p.add("linear", C.byref(a1))
        p.add("rmsnorm", h.data_ptr(), H, nw.data_ptr(), xn.data_ptr(), H, T, H, 1e-5)
        p.add("linear", C.byref(a2))
        p.add("linear", C.byref(a3))

    p = slab_program(build, [(w1, L.EPI_STORE), (w3, L.EPI_STORE)])
    assert p.phase_types() == [L.PH_GEMM2, L.PH_NORM, L.PH_GEMM, L.PH_EPI, L.PH_GEMM2]
    assert_bf16_close(h[:T], (x.float() @ w1.float().T).to(BF))
    hf = h[:T].float()
    xn_ref = nw * (hf * torch.rsqrt(hf.pow(2).mean(-1, keepdim=True) + 1e-5)).to(BF)
    assert_bf16_close(xn[:T], xn_ref)
    assert_bf16_close(y[:T], (xn[:T].float() @ w2.float().T).to(BF))
    assert_bf16_close(z, (y[:T].float() @ w3.float().T).to(BF))
    for _ in range(3):
        z1 = z.clone()
        p.run()
        torch.cuda.synchronize()
        assert torch.equal(z, z1)


@pytest.mark.parametrize("N,K,epi", [(4096, 4096, L.EPI_RESID), (12288, 4096, L.EPI_HEADS), (22016, 4096, L.EPI_SWIGLU),
                                     (32064, 4096, L.EPI_STORE), (1000, 256, L.EPI_STORE)])
def test_slab_pack_kernel_matches_the_python_restatement(N, K, epi):
    """vs_slab_pack (CUDA) against tests/slab_model.py (an independent restatement of csrc/slab.h), byte for byte, on a
    slice of k-chunks (the restatement is a Python loop)."""
    import numpy as np
    import slab_model as sm
    from vispec_b200.engine import pack_weight_slab
    Ks = min(K, 256)
    w = randn(N, Ks, seed=71)
    slab = pack_weight_slab(w, epi)
    p = sm.plan(L.load(), N, Ks, epi, 0)
    ref = sm.pack(p, w.view(torch.int16).cpu().numpy().view(np.uint16))
    got = slab.view(torch.int16).cpu().numpy().view(np.uint16)
    assert got.shape == ref.shape and (got == ref).all()
