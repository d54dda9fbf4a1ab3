"""Kernel-level parity (GPU): every C-ABI op against a plain fp32 PyTorch statement of the same op with the
reference's bf16 rounding points.  Tolerances are stated per test; integer outputs are compared exactly."""
import ctypes as C
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from vispec_b200 import _lib as L  # noqa: E402
from vispec_b200.engine import ScratchBuffers, pack_gate_up, rope_tables  # noqa: E402

DEV = "cuda"
BF = torch.bfloat16


def stream():
    return torch.cuda.current_stream().cuda_stream


def rbf(x):
    return x.to(BF).float()


def randn(*s, std=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*s, generator=g) * std).to(BF).to(DEV)


def linear(x, w, bias, T, N, K, epi=L.EPI_STORE, y=None, ldy=None, R=None, splits=None, **kw):
    lib = L.load()
    sp = lib.vs_linear_splits(N, K, T) if splits is None else splits
    partials = torch.zeros(max(1, lib.vs_linear_partials_floats(N, T, sp)), dtype=torch.float32, device=DEV)
    counters = torch.zeros(4096, dtype=torch.int32, device=DEV)
    a = L.LinearArgs()
    a.X, a.ldx, a.W, a.ldw, a.bias = x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0), L.ptr(bias)
    a.T, a.N, a.K, a.splits = T, N, K, sp
    a.partials, a.counters, a.epi = partials.data_ptr(), counters.data_ptr(), epi
    if y is not None:
        a.Y, a.ldy = y.data_ptr(), ldy if ldy is not None else y.stride(0)
    if R is not None:
        a.R, a.ldr = R.data_ptr(), R.stride(0)
    for k, v in kw.items():
        setattr(a, k, v)
    L.check(lib.vs_linear(C.byref(a), stream()), "vs_linear")
    torch.cuda.synchronize()
    assert int(counters.abs().sum()) == 0, "split-K counters must be left at zero"
    return sp


# fp32-accumulated reference: differences come only from accumulation order, so outputs agree to 1 bf16 ulp
# on a small fraction of elements; we allow rtol 2^-7 (one ulp) and require >=99% exact matches.
def assert_bf16_close(got, ref, frac_exact=0.99, atol=1e-2):
    got, ref = got.float(), ref.float()
    assert torch.isfinite(got).all()
    torch.testing.assert_close(got, ref, rtol=2 ** -7, atol=atol)
    exact = (got == ref).float().mean().item()
    assert exact >= frac_exact, f"only {exact:.4f} of elements bit-exact"


@pytest.mark.parametrize("T,N,K", [(1, 256, 256), (5, 1000, 256), (30, 4096, 4096), (32, 4096, 11008), (8, 32064, 4096),
                                   (45, 768, 512), (30, 4096, 8192),
                                   # more than 64 rows: the prompt-pass kernel (gemm_wide.cu, 256-token MMA tiles)
                                   (65, 256, 256), (129, 4096, 4096), (256, 1000, 512), (257, 384, 4096), (704, 4096, 11008),
                                   (576, 4096, 4096)])
@pytest.mark.parametrize("bias", [False, True])
def test_linear_store(T, N, K, bias):
    x, w = randn(T, K, seed=1), randn(N, K, std=K ** -0.5, seed=2)
    b = randn(N, std=0.1, seed=3) if bias else None
    y = torch.zeros(T, N, dtype=BF, device=DEV)
    linear(x, w, b, T, N, K, y=y)
    ref = x.float() @ w.float().T
    if bias:
        ref = ref + b.float()
    assert_bf16_close(y, ref.to(BF))


@pytest.mark.parametrize("splits", [1, 2, 5])
def test_linear_split_k_is_deterministic(splits):
    T, N, K = 30, 1024, 4096
    x, w = randn(T, K, seed=4), randn(N, K, std=K ** -0.5, seed=5)
    outs = []
    for _ in range(3):
        y = torch.zeros(T, N, dtype=BF, device=DEV)
        linear(x, w, None, T, N, K, y=y, splits=splits)
        outs.append(y)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    assert_bf16_close(outs[0], (x.float() @ w.float().T).to(BF))


@pytest.mark.parametrize("T", [30, 129, 704])
def test_linear_residual(T):
    N, K = 4096, 4096
    x, w, r = randn(T, K, seed=6), randn(N, K, std=K ** -0.5, seed=7), randn(T, N, seed=8)
    y = r.clone()
    linear(x, w, None, T, N, K, epi=L.EPI_RESID, y=y, R=y)          # in place, as the decoder layer uses it
    ref = (r.float() + rbf(x.float() @ w.float().T)).to(BF)
    # one bf16 ulp of the linear output (|acc| < 4: 2^-6) survives a cancelling residual add unchanged
    assert torch.isfinite(y.float()).all()
    torch.testing.assert_close(y.float(), ref.float(), rtol=2 ** -7, atol=2 ** -6)
    assert (y == ref).float().mean().item() >= 0.99


@pytest.mark.parametrize("T", [30, 300])
def test_linear_swiglu(T):
    H, I = 4096, 11008
    x = randn(T, H, seed=9)
    gate, up = randn(I, H, std=H ** -0.5, seed=10), randn(I, H, std=H ** -0.5, seed=11)
    w = pack_gate_up(gate, up)
    y = torch.zeros(T, I, dtype=BF, device=DEV)
    linear(x, w, None, T, 2 * I, H, epi=L.EPI_SWIGLU, y=y)
    g, u = rbf(x.float() @ gate.float().T), rbf(x.float() @ up.float().T)
    ref = (rbf(torch.nn.functional.silu(g)) * u).to(BF)
    # silu and the product amplify 1-ulp differences of g and u: a few bf16 ulps at single outputs, >= 97 % bit-exact
    torch.testing.assert_close(y.float(), ref.float(), rtol=2 ** -5, atol=2e-2)
    assert (y == ref).float().mean().item() >= 0.97


@pytest.mark.parametrize("nh,nkv,T", [(32, 32, 30), (4, 2, 7), (2, 2, 40), (4, 2, 200), (8, 8, 704)])
def test_linear_heads_rope_and_kv_append(nh, nkv, T):
    H, D, max_rows, max_pos = nh * 128, 128, 1024, 1024
    x = randn(T, H, seed=12)
    w = randn((nh + 2 * nkv) * D, H, std=H ** -0.5, seed=13)
    b = randn((nh + 2 * nkv) * D, std=0.1, seed=14)
    cos, sin = rope_tables(D, max_pos, 10000.0, BF, DEV)
    pos = torch.randint(0, max_pos, (T,), dtype=torch.int32, device=DEV)
    base = torch.tensor([17], dtype=torch.int32, device=DEV)
    q_rows = (T + 31) // 32 * 32
    Q = torch.zeros(nh, q_rows, D, dtype=BF, device=DEV)
    Kc = torch.zeros(nkv, max_rows, D, dtype=BF, device=DEV)
    Vc = torch.zeros(nkv, max_rows, D, dtype=BF, device=DEV)
    linear(x, w, b, T, (nh + 2 * nkv) * D, H, epi=L.EPI_HEADS, n_q_tiles=nh, n_k_tiles=nkv, rope=1,
           cos=cos.data_ptr(), sin=sin.data_ptr(), pos_ids=pos.data_ptr(), Qout=Q.data_ptr(), q_head_stride=q_rows * D,
           Kout=Kc.data_ptr(), Vout=Vc.data_ptr(), kv_head_stride=max_rows * D, kv_row_base=base.data_ptr(), kv_row_off=3)
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


def test_rmsnorm():
    T, H = 30, 4096
    x, w = randn(T, H, seed=15), (1 + 0.1 * torch.randn(H)).to(BF).to(DEV)
    y = torch.zeros_like(x)
    L.check(L.load().vs_rmsnorm(x.data_ptr(), H, w.data_ptr(), y.data_ptr(), H, T, H, 1e-5, stream()))
    xf = x.float()
    ref = w * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5)).to(BF)
    assert_bf16_close(y, ref)


def test_gather_embed_broadcast():
    lib = L.load()
    tab = randn(1000, 256, seed=16)
    idx = torch.tensor([5, 999, 0, 5], dtype=torch.int32, device=DEV)
    dst = torch.zeros(4, 512, dtype=BF, device=DEV)
    L.check(lib.vs_embed_rows(tab.data_ptr(), 256, idx.data_ptr(), dst.data_ptr(), 512, 4, 256, 1000, stream()))
    assert torch.equal(dst[:, :256], tab[idx.long()]) and int(dst[:, 256:].abs().sum()) == 0
    L.check(lib.vs_broadcast_row(tab[7].data_ptr(), dst[:, 256:].data_ptr(), 512, 4, 256, stream()))
    assert torch.equal(dst[:, 256:], tab[7].expand(4, 256))
    out = torch.zeros(3, 256, dtype=BF, device=DEV)
    L.check(lib.vs_gather_rows(dst.data_ptr(), 512, None, out.data_ptr(), 256, 3, 256, stream()))
    assert torch.equal(out, dst[:3, :256])
    pick = torch.tensor([3, 0, 3], dtype=torch.int32, device=DEV)
    L.check(lib.vs_gather_rows(dst.data_ptr(), 512, pick.data_ptr(), out.data_ptr(), 256, 3, 256, stream()))
    assert torch.equal(out, dst[pick.long(), :256])


def test_argmax_first_index_on_ties():
    V = 32064
    x = randn(30, V, std=4.0, seed=17)
    x[3, 100] = x[3, 20000] = 50.0          # exact tie: the lower index wins (torch.argmax rule)
    out = torch.zeros(30, dtype=torch.int32, device=DEV)
    L.check(L.load().vs_argmax_rows(x.data_ptr(), V, 30, V, out.data_ptr(), stream()))
    ref = torch.argmax(x.float(), dim=-1)
    assert int(out[3]) == 100
    assert torch.equal(out.long(), ref)


def ref_topk_lowest_index(vals, k):
    """(value desc, index asc): the library's tie rule."""
    v, i = torch.sort(vals, dim=-1, descending=True, stable=True)
    return v[..., :k], i[..., :k]


@pytest.mark.parametrize("V", [1000, 32064])
def test_logsoftmax_topk(V):
    T, k = 8, 8
    x = randn(T, V, std=3.0, seed=18)
    tv = torch.zeros(T * k, dtype=torch.float32, device=DEV)
    ti = torch.zeros(T * k, dtype=torch.int32, device=DEV)
    L.check(L.load().vs_logsoftmax_topk(x.data_ptr(), V, T, V, k, None, tv.data_ptr(), ti.data_ptr(), stream()))
    lp = torch.log_softmax(x.float(), dim=-1).to(BF).float()      # fp32 math, bf16 result (cnets_ours.py:1111-1113)
    rv, ri = ref_topk_lowest_index(lp, k)
    got_v, got_i = tv.view(T, k), ti.view(T, k).long()
    # values: same bf16 number up to one ulp of the log-sum-exp; indices exact wherever the reference values are
    # separated by more than that ulp
    torch.testing.assert_close(got_v, rv, rtol=0, atol=2 ** -5)
    sep = (lp.gather(1, got_i) - got_v).abs().max().item()
    assert sep <= 2 ** -5
    same = (got_i == ri)
    gaps = (rv[:, :-1] - rv[:, 1:]).abs().min(dim=1).values
    for r in range(T):
        if gaps[r] > 2 ** -4:
            assert same[r].all(), f"row {r}: {got_i[r].tolist()} vs {ri[r].tolist()}"
    # single-row selection through a device index
    sel = torch.tensor([5], dtype=torch.int32, device=DEV)
    tv1 = torch.zeros(k, dtype=torch.float32, device=DEV)
    ti1 = torch.zeros(k, dtype=torch.int32, device=DEV)
    L.check(L.load().vs_logsoftmax_topk(x.data_ptr(), V, 1, V, k, sel.data_ptr(), tv1.data_ptr(), ti1.data_ptr(), stream()))
    assert torch.equal(tv1, got_v[5]) and torch.equal(ti1.long(), got_i[5])


def test_logsoftmax_topk_tie_rule():
    V, k = 1000, 8
    x = torch.full((1, V), -2.0, dtype=BF, device=DEV)
    x[0, [900, 17, 300]] = 5.0
    tv = torch.zeros(k, dtype=torch.float32, device=DEV)
    ti = torch.zeros(k, dtype=torch.int32, device=DEV)
    L.check(L.load().vs_logsoftmax_topk(x.data_ptr(), V, 1, V, k, None, tv.data_ptr(), ti.data_ptr(), stream()))
    assert ti.tolist() == [17, 300, 900, 0, 1, 2, 3, 4]


# ---- attention --------------------------------------------------------------------------------------------
def attention_ref(Q, K, V, n_ctx, n_new, causal, bits, eager, nh, nkv):
    """Q [nh,Tq,128]; K,V [nkv,rows,128] -> [Tq, nh*128] in the two rounding regimes."""
    Tq, S, D = Q.shape[1], n_ctx + n_new, 128
    rep = nh // nkv
    Kf = K[:, :S].float().repeat_interleave(rep, dim=0)
    Vf = V[:, :S].float().repeat_interleave(rep, dim=0)
    vis = torch.zeros(Tq, S, dtype=torch.bool, device=Q.device)
    vis[:, :n_ctx] = True
    for i in range(Tq):
        if causal:
            vis[i, :min(S, n_ctx + i + 1)] = True
        else:
            for j in range(min(n_new, 64)):
                if (int(bits[i]) >> j) & 1:
                    vis[i, n_ctx + j] = True
    s = torch.einsum("hqd,hkd->hqk", Q.float(), Kf)
    if eager:
        s = rbf(rbf(s) / math.sqrt(D))
    else:
        s = s / math.sqrt(D)
    s = s.masked_fill(~vis[None], float("-inf"))
    if eager:
        p = rbf(torch.softmax(s, dim=-1))
        o = torch.einsum("hqk,hkd->hqd", p, Vf)
    else:
        m = s.max(dim=-1, keepdim=True).values
        e = torch.exp(s - m)
        o = torch.einsum("hqk,hkd->hqd", rbf(e), Vf) / e.sum(-1, keepdim=True)
    return o.transpose(0, 1).reshape(Tq, nh * D).to(BF)


def run_attention(Q, K, V, nh, nkv, Tq, n_ctx, n_new, causal, bits, eager, n_ctx_dev=True):
    rows_q = Q.shape[1]
    sc = ScratchBuffers(nh * 128, 256, nh, 1000, rows_q, K.shape[1], DEV)
    O = torch.zeros(Tq, nh * 128, dtype=BF, device=DEV)
    a = L.AttnArgs()
    a.Q, a.q_head_stride, a.K, a.V, a.kv_head_stride = Q.data_ptr(), rows_q * 128, K.data_ptr(), V.data_ptr(), K.shape[1] * 128
    a.n_heads, a.n_kv_heads, a.Tq = nh, nkv, Tq
    nctx_t = torch.tensor([n_ctx], dtype=torch.int32, device=DEV)
    a.n_ctx, a.n_ctx_host = (nctx_t.data_ptr(), 0) if n_ctx_dev else (None, n_ctx)
    a.n_new, a.causal, a.bits, a.eager, a.scale = n_new, int(causal), L.ptr(bits), int(eager), 1 / math.sqrt(128)
    a.scores, a.score_ld, a.stats, a.max_chunks = sc.scores.data_ptr(), sc.score_ld, sc.stats.data_ptr(), sc.max_chunks
    a.opart, a.pv_splits, a.counters, a.O, a.ldo = sc.opart.data_ptr(), sc.pv_splits, sc.attn_counters.data_ptr(), O.data_ptr(), nh * 128
    L.check(L.load().vs_attention(C.byref(a), stream()), "vs_attention")
    torch.cuda.synchronize()
    assert int(sc.attn_counters.abs().sum()) == 0
    return O


def tree_bits(T, seed):
    """random ancestor-closed tree masks like cnets_ours.py:1183-1186."""
    g = torch.Generator().manual_seed(seed)
    bits = [1]
    for i in range(1, T):
        parent = int(torch.randint(0, i, (1,), generator=g))
        bits.append(bits[parent] | (1 << i))
    return torch.tensor(bits, dtype=torch.int64, device=DEV)


@pytest.mark.parametrize("eager", [True, False])
@pytest.mark.parametrize("nh,nkv,n_ctx", [(32, 32, 704), (4, 2, 61), (2, 2, 0), (8, 8, 3000)])
def test_attention_tree(eager, nh, nkv, n_ctx):
    T, rows = 30, 4096
    Q = randn(nh, 32, 128, seed=20)
    K, V = randn(nkv, rows, 128, seed=21), randn(nkv, rows, 128, seed=22)
    bits = tree_bits(T, 5)
    O = run_attention(Q, K, V, nh, nkv, T, n_ctx, T, False, bits, eager)
    ref = attention_ref(Q[:, :T], K, V, n_ctx, T, False, bits.tolist(), eager, nh, nkv)
    # softmax weights are rounded to bf16 in both; a different fp32 summation order moves single weights by an ulp
    torch.testing.assert_close(O.float(), ref.float(), rtol=2 ** -6, atol=2e-2)
    assert (O.float() == ref.float()).float().mean() > 0.9


@pytest.mark.parametrize("Tq,n_ctx", [(1, 100), (5, 129), (8, 40), (70, 0), (129, 0), (100, 300), (704, 0), (33, 900)])
def test_attention_causal(Tq, n_ctx):
    nh = nkv = 4
    rows_q = (Tq + 31) // 32 * 32
    Q = randn(nh, rows_q, 128, seed=23)
    K, V = randn(nkv, 1024, 128, seed=24), randn(nkv, 1024, 128, seed=25)
    O = run_attention(Q, K, V, nh, nkv, Tq, n_ctx, Tq, True, None, False, n_ctx_dev=(n_ctx > 0))
    ref = attention_ref(Q[:, :Tq], K, V, n_ctx, Tq, True, None, False, nh, nkv)
    torch.testing.assert_close(O.float(), ref.float(), rtol=2 ** -6, atol=2e-2)


def test_attention_cross_no_mask():
    """ImgAdaptor: num_q=2 learned queries over 576 image keys, nothing masked (cnets_ours.py:649)."""
    nh = 32
    Q = randn(nh, 32, 128, std=128 ** -0.5, seed=26)
    K, V = randn(nh, 576, 128, seed=27), randn(nh, 576, 128, seed=28)
    O = run_attention(Q, K, V, nh, nh, 2, 576, 0, False, None, False, n_ctx_dev=False)
    ref = attention_ref(Q[:, :2], K, V, 576, 0, False, [0, 0], False, nh, nh)
    torch.testing.assert_close(O.float(), ref.float(), rtol=2 ** -6, atol=2e-2)


# ---- KV splice ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("sel", [[0], [0, 3, 9, 29], [0, 1, 2], [0, 2, 3, 4, 28]])
def test_kv_splice_matches_gather_then_copy(sel):
    n_slabs, n_kv, max_pos, prev = 6, 4, 256, 100
    slab = randn(n_slabs, 1, n_kv, max_pos, 128, seed=29)
    ref = slab.clone()
    idx = torch.tensor(sel, device=DEV) + prev
    tgt = ref[..., idx, :]
    ref[..., prev:prev + len(sel), :] = tgt                                  # utils.py:529-541
    host = (C.c_int32 * L.VS_MAX_ACCEPT)(*sel)
    L.check(L.load().vs_kv_splice(slab.data_ptr(), n_slabs, n_kv, max_pos, None, None, prev, len(sel), host, stream()))
    torch.cuda.synchronize()
    assert torch.equal(slab, ref)
    # device-driven variant
    slab2 = randn(n_slabs, 1, n_kv, max_pos, 128, seed=29)
    st = torch.zeros(L.ST_SLOTS, dtype=torch.int32, device=DEV)
    st[L.ST_PREV_LEN], st[L.ST_ACCEPT] = prev, len(sel) - 1
    sel_dev = torch.tensor(sel + [sel[-1]] * (L.VS_MAX_ACCEPT - len(sel)), dtype=torch.int32, device=DEV)
    L.check(L.load().vs_kv_splice(slab2.data_ptr(), n_slabs, n_kv, max_pos, st.data_ptr(), sel_dev.data_ptr(), 0, 0, None, stream()))
    torch.cuda.synchronize()
    assert torch.equal(slab2, ref)
