"""Integer kernels (GPU) against the oracle: draft-tree bookkeeping, tree build (A7), greedy accept (A8).
Bit-exact on tie-free inputs; the tie rule itself (value desc, index asc) is tested separately."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import vispec_oracle as vo  # noqa: E402  (checker only)
from vispec_b200 import _lib as L  # noqa: E402
from vispec_b200.engine import DraftConfig  # noqa: E402

DEV = "cuda"


def stream():
    return torch.cuda.current_stream().cuda_stream


class TreeState:
    """Device buffers of vs_tree_bufs without a model around them."""

    def __init__(self, top_k=8, depth=3, total_tokens=30):
        k, total = top_k, total_tokens - 1
        n_sc = k + depth * k * k
        z = lambda n, dt: torch.zeros(n, dtype=dt, device=DEV)
        i32, f32, i64 = torch.int32, torch.float32, torch.int64
        self.t = dict(scores_flat=z(n_sc, f32), tokens_flat=z(n_sc, i32), parents=z(1 + depth * k, i32),
                      frontier_scores=z(k, f32), frontier_ids=z(32, i32), cs_index=z(k, i32), frontier_bits=z(32, i64),
                      top_val=z(k * k, f32), top_idx=z(k * k, i32), gather_idx=z(32, i32), draft_pos_ids=z(32, i32),
                      draft_tokens=z(32, i32), tree_bits=z(32, i64), tree_pos=z(32, i32), verify_pos=z(32, i32),
                      retrieve=z(32 * L.VS_RETRIEVE_LD, i32), sel=z(L.VS_MAX_ACCEPT, i32), stable_ids=z(32, i32),
                      stable_pos=z(32, i32))
        self.tb = L.TreeBufs(k, depth, total, *[self.t[n].data_ptr() for n, _ in L.TreeBufs._fields_[3:]])
        self.state = z(L.ST_SLOTS, i32)
        self.k, self.depth, self.total = k, depth, total

    def outputs(self):
        n = self.total + 1
        st = self.state.tolist()
        nl, md = st[L.ST_N_LEAVES], st[L.ST_MAX_DEPTH]
        bits = self.t["tree_bits"][:n]
        mask = ((bits[:, None] >> torch.arange(n, device=DEV)[None]) & 1).float()[None, None].cpu()
        return (self.t["draft_tokens"][:n].long().cpu()[None], self.t["retrieve"].view(32, -1)[:nl, :md].long().cpu(),
                mask, self.t["tree_pos"][:n].long().cpu())


def simulate_expansion(seed, k=8, depth=3, vocab=1000, unique=True):
    """The bookkeeping of cnets_ours.py:1109-1165 on random per-level top-k log-probs (no model), via the same
    statements the oracle uses.  Returns per-level inputs and the final (scores_flat, tokens_flat, parents)."""
    g = torch.Generator().manual_seed(seed)

    def level_vals(rows):
        v = -torch.rand(rows, k, generator=g) * 6
        if not unique:
            v = (v * 2).round() / 2                 # many exact ties
        v = v.to(torch.bfloat16).float().sort(dim=-1, descending=True).values
        idx = torch.stack([torch.randperm(vocab, generator=g)[:k] for _ in range(rows)])
        return v, idx

    levels = []
    v0, i0 = level_vals(1)
    scores = v0[0]
    scores_list, tokens_list, parents_list = [scores[None]], [i0], [torch.zeros(1, dtype=torch.long)]
    cs_index = torch.arange(k)
    tree_mask = torch.eye(k)[None, None]
    for i in range(depth):
        tv, ti = level_vals(k)
        bias = 1 + k * k * max(0, i - 1) + (k if i > 0 else 0)
        parents_list.append(cs_index + bias)
        cu = (tv + scores[:, None]).to(torch.bfloat16).float()
        srt = torch.sort(cu.view(-1), descending=True, stable=True)        # (value desc, index asc)
        cs_index, scores = srt.indices[:k], srt.values[:k]
        out_ids = cs_index // k
        tree_mask = torch.cat((tree_mask[:, :, out_ids], torch.eye(k)[None, None]), dim=3)
        levels.append(dict(top_val=tv, top_idx=ti, cs_index=cs_index.clone(), scores=scores.clone(),
                           ids=ti.view(-1)[cs_index].clone(), out_ids=out_ids.clone(), mask=tree_mask.clone()))
        scores_list.append(cu)
        tokens_list.append(ti)
    return dict(level0=(v0[0], i0[0]), levels=levels, scores_flat=torch.cat(scores_list, 0).view(-1),
                tokens_flat=torch.cat(tokens_list, 0).view(-1), parents=torch.cat(parents_list, 0))


@pytest.mark.parametrize("seed", range(6))
def test_draft_expansion_bookkeeping(seed):
    lib = L.load()
    ts = TreeState()
    sim = simulate_expansion(seed)
    k = ts.k
    ts.state[L.ST_D_POS] = 77
    ts.state[L.ST_LAST_ROW] = 2
    ts.t["top_val"][:k] = sim["level0"][0].to(DEV)
    ts.t["top_idx"][:k] = sim["level0"][1].to(DEV, torch.int32)
    L.check(lib.vs_draft_level0(C.byref(ts.tb), ts.state.data_ptr(), stream()))
    assert ts.t["gather_idx"][:k].tolist() == [2] * k
    assert ts.t["frontier_bits"][:k].tolist() == [1 << j for j in range(k)]
    for i, lv in enumerate(sim["levels"]):
        ts.t["top_val"].copy_(lv["top_val"].view(-1).to(DEV))
        ts.t["top_idx"].copy_(lv["top_idx"].view(-1).to(DEV, torch.int32))
        L.check(lib.vs_draft_expand(C.byref(ts.tb), ts.state.data_ptr(), i, stream()))
        assert ts.t["cs_index"].tolist() == lv["cs_index"].tolist()
        assert torch.equal(ts.t["frontier_scores"].cpu(), lv["scores"])
        assert ts.t["frontier_ids"][:k].tolist() == lv["ids"].tolist()
        assert ts.t["gather_idx"][:k].tolist() == lv["out_ids"].tolist()
        assert ts.t["draft_pos_ids"][:k].tolist() == [77 + i + 1] * k
        m = lv["mask"][0, 0]                                       # [k, k*(i+2)]
        want = [(int(sum(int(m[r, c]) << c for c in range(m.shape[1])))) for r in range(k)]
        assert ts.t["frontier_bits"][:k].tolist() == want
    assert torch.equal(ts.t["scores_flat"].cpu(), sim["scores_flat"])
    assert ts.t["tokens_flat"].tolist() == sim["tokens_flat"].tolist()
    assert ts.t["parents"].tolist() == sim["parents"].tolist()


@pytest.mark.parametrize("seed", range(8))
@pytest.mark.parametrize("shape", [(8, 3, 30), (8, 5, 32), (4, 2, 10)])
def test_tree_build_matches_oracle(seed, shape):
    k, depth, total_tokens = shape
    lib = L.load()
    ts = TreeState(k, depth, total_tokens)
    sim = simulate_expansion(seed, k=k, depth=depth)
    # tie-free selection scores (unique fp32 values); tree_build only ranks them
    g = torch.Generator().manual_seed(100 + seed)
    scores = sim["scores_flat"] + torch.randperm(sim["scores_flat"].numel(), generator=g).float() * 2e-5
    assert scores.unique().numel() == scores.numel()
    ts.t["scores_flat"].copy_(scores.to(DEV))
    ts.t["tokens_flat"].copy_(sim["tokens_flat"].to(DEV, torch.int32))
    ts.t["parents"].copy_(sim["parents"].to(DEV, torch.int32))
    ts.state[L.ST_CUR_LEN] = 500
    L.check(lib.vs_tree_build(C.byref(ts.tb), ts.state.data_ptr(), None, 4242, stream()))
    got = ts.outputs()
    ref = vo.tree_build(scores, sim["tokens_flat"], sim["parents"], torch.tensor([4242]), ts.total, k)
    assert torch.equal(got[0], ref[0])
    assert torch.equal(got[1], ref[1])
    assert torch.equal(got[2], ref[2])
    assert torch.equal(got[3], ref[3])
    assert ts.t["verify_pos"][:ts.total + 1].tolist() == (ref[3] + 500).tolist()


def test_tree_build_tie_rule_keeps_parents():
    """With exact score ties (log-prob 0 children, as in the bigram fixture) the lower flat index wins, so a
    parent is always selected before its child and every parent lookup is valid."""
    lib = L.load()
    ts = TreeState()
    sim = simulate_expansion(3, unique=False)
    ts.t["scores_flat"].copy_(sim["scores_flat"].to(DEV))
    ts.t["tokens_flat"].copy_(sim["tokens_flat"].to(DEV, torch.int32))
    ts.t["parents"].copy_(sim["parents"].to(DEV, torch.int32))
    L.check(lib.vs_tree_build(C.byref(ts.tb), ts.state.data_ptr(), None, 1, stream()))
    sc = sim["scores_flat"]
    order = torch.sort(sc, descending=True, stable=True).indices[:ts.total].sort().values      # the documented rule
    ref = vo.tree_build(torch.zeros_like(sc).index_fill(0, order, 1.0) + torch.arange(sc.numel()).flip(0) * 1e-6,
                        sim["tokens_flat"], sim["parents"], torch.tensor([1]), ts.total, 8)
    got = ts.outputs()
    for a, b in zip(got, ref):
        assert torch.equal(a, b)


@pytest.mark.parametrize("seed", range(10))
def test_accept_greedy_matches_oracle(seed):
    lib = L.load()
    ts = TreeState()
    sim = simulate_expansion(seed)
    g = torch.Generator().manual_seed(200 + seed)
    ts.t["scores_flat"].copy_((sim["scores_flat"] + torch.randperm(200, generator=g).float() * 2e-5).to(DEV))
    # small vocabulary so that draft tokens match the target's arg-max often (accept lengths 0..depth+1)
    tokens = torch.randint(0, 3, (200,), generator=g)
    ts.t["tokens_flat"].copy_(tokens.to(DEV, torch.int32))
    ts.t["parents"].copy_(sim["parents"].to(DEV, torch.int32))
    ts.state[L.ST_CUR_LEN] = 300
    ts.state[L.ST_D_POS] = 300
    L.check(lib.vs_tree_build(C.byref(ts.tb), ts.state.data_ptr(), None, 2, stream()))
    draft_tokens, retrieve, _, _ = ts.outputs()
    n = ts.total + 1
    amax = torch.randint(0, 3, (n,), generator=g)
    amax_dev = torch.zeros(32, dtype=torch.int32, device=DEV)
    amax_dev[:n] = amax.to(DEV)
    rec = torch.zeros(2 + L.VS_MAX_ACCEPT, dtype=torch.int32, device=DEV)
    L.check(lib.vs_accept_greedy(C.byref(ts.tb), ts.state.data_ptr(), amax_dev.data_ptr(), rec.data_ptr(), stream()))
    # oracle: one-hot logits per tree slot, gathered along the paths like utils.py:411
    node_logits = torch.nn.functional.one_hot(amax, 3).float()
    logits = node_logits[retrieve]
    cand = torch.cat((draft_tokens, torch.full((1, 1), -1)), dim=1)[0, retrieve]
    best, a, sample_p = vo.evaluate_posterior_greedy(logits, cand)
    st = ts.state.tolist()
    assert (st[L.ST_BEST], st[L.ST_ACCEPT]) == (best, a)
    assert st[L.ST_NEXT_TOKEN] == int(torch.argmax(sample_p))
    assert st[L.ST_PREV_LEN] == 300 and st[L.ST_CUR_LEN] == 300 + a + 1 and st[L.ST_M] == a + 1
    assert ts.t["sel"][:a + 1].tolist() == retrieve[best, :a + 1].tolist()
    r = rec.tolist()
    assert r[0] == a and r[1] == st[L.ST_NEXT_TOKEN] and r[2:3 + a] == cand[best, :a + 1].tolist()
    want_ids = cand[best, 1:a + 1].tolist() + [st[L.ST_NEXT_TOKEN]]
    assert ts.t["stable_ids"][:a + 1].tolist() == want_ids
    assert ts.t["stable_pos"][:a + 1].tolist() == list(range(300, 301 + a))
