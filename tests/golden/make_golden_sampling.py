#!/usr/bin/env python
"""Golden vectors for the T > 0 acceptance from the UNMODIFIED reference.

Run in the build container only:   python tests/golden/make_golden_sampling.py

Drives the reference's own `vispec.model.utils.evaluate_posterior(logits, candidates, logits_processor)` with the
`prepare_logits_processor(temperature=T)` it builds for sampling, on random draft trees (built by the reference-pinned
oracle tree build with the T > 0 path sorting) and random peaked tree logits.  The only intervention is that
`torch.rand_like` is replaced, for the duration of the call, by a function returning a preset uniform tensor, so the
accept / reject decisions are reproducible; what the reference returned is saved to tests/golden/posterior_sampling.pt.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
import transformers.cache_utils as _cu  # noqa: E402

if not hasattr(_cu, "SlidingWindowCache"):
    _cu.SlidingWindowCache = type("SlidingWindowCache", (), {})
from vispec.model import utils as ref_utils  # noqa: E402

from oracle import vispec_oracle as vo  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def random_tree(g, V, top_k=8, depth=3, total=29):
    """A draft tree shaped like topK_genrate's: random cumulative scores over top_k + depth*top_k^2 nodes."""
    n_sc = top_k + depth * top_k * top_k
    scores = (torch.randn(n_sc, generator=g) * 2).to(torch.bfloat16).float()
    tokens = torch.randint(0, V, (n_sc,), generator=g)
    # children of one parent must be distinct tokens (they are a top-k of one distribution)
    for s in range(0, n_sc, top_k):
        tokens[s:s + top_k] = torch.randperm(V, generator=g)[:top_k]
    # parents: group 0 hangs off the root; later groups off random nodes of the previous level
    parents = [0]
    for lvl in range(depth):
        lo = 1 if lvl == 0 else 1 + top_k + (lvl - 1) * top_k * top_k
        hi = lo + (top_k if lvl == 0 else top_k * top_k)
        parents += sorted(torch.randperm(hi - lo, generator=g)[:top_k].add(lo).tolist())
    parents = torch.tensor(parents)
    root = torch.randint(0, V, (1,), generator=g)
    return vo.tree_build(scores, tokens, parents, root, total, top_k, sort_paths=True)


def main():
    g = torch.Generator().manual_seed(1234)
    V = 512
    cases = []
    for ci in range(32):
        temperature = [0.5, 1.0, 1.5, 2.0][ci % 4]
        draft_tokens, retrieve, _, _ = random_tree(g, V)
        T = draft_tokens.shape[1]
        tree_logits = (torch.randn(T, V, generator=g) * 2.0).to(torch.bfloat16).float()
        # make some children likely under their parent's distribution so that accepts happen
        for leaf in retrieve.tolist():
            for d in range(1, len(leaf)):
                if leaf[d] >= 0 and torch.rand(1, generator=g).item() < 0.6:
                    tree_logits[leaf[d - 1], draft_tokens[0, leaf[d]]] += 6.0 * torch.rand(1, generator=g).item()
        tree_logits = tree_logits.to(torch.bfloat16).float()
        logits = tree_logits[retrieve]                                   # utils.py:411 (negative index = last row)
        pad = torch.full((1, 1), -1, dtype=torch.long)
        candidates = torch.cat((draft_tokens, pad), dim=1)[0, retrieve]
        uni = torch.rand(candidates.shape, generator=g)
        proc = ref_utils.prepare_logits_processor(temperature=temperature)
        orig = torch.rand_like
        torch.rand_like = lambda t, dtype=None, **kw: uni.to(dtype or t.dtype)
        try:
            best, acc, sample_p = ref_utils.evaluate_posterior(logits.clone(), candidates.clone(), proc)
        finally:
            torch.rand_like = orig
        cases.append(dict(temperature=temperature, draft_tokens=draft_tokens, retrieve=retrieve,
                          tree_logits=tree_logits.to(torch.bfloat16),      # exact: the values are bf16-representable
                          uni=uni, best=int(best), accept_length=int(acc), sample_p=sample_p.clone()))
    torch.save(dict(cases=cases, V=V, torch=torch.__version__), os.path.join(OUT, "posterior_sampling.pt"))
    print("accept lengths:", [c["accept_length"] for c in cases])


if __name__ == "__main__":
    main()
