"""Golden vectors for oracle.vispec_oracle.warp_logits from the UNMODIFIED reference's `prepare_logits_processor`
(vispec/model/utils.py:39-55; the warpers themselves are transformers'), plus the observation that the reference's own
nucleus filter cannot run on the 3-D tree logits of `evaluate_posterior` (utils.py:454).
    python tests/golden/make_golden_warpers.py      (build container; needs /root/reference)"""
import os
import sys

import torch

torch.cuda.synchronize = lambda *a, **k: None      # caller-side shim: importing utils builds nothing on the GPU, Timer would
sys.path.insert(0, "/root/reference")
from vispec.model.utils import prepare_logits_processor  # noqa: E402

g = torch.Generator().manual_seed(7)
x = torch.randn(9, 517, generator=g) * 3.0
x[3, 100:110] = x[3, 100]                            # a run of exact ties
cases = []
for kw in [dict(temperature=0.7, top_p=0.9), dict(temperature=1.0, top_p=0.5), dict(temperature=0.7, top_k=20),
           dict(temperature=1.3, top_p=0.8, top_k=5), dict(temperature=0.5), dict(temperature=1.0, top_k=1),
           dict(temperature=2.0, top_p=0.05)]:
    cases.append(dict(kw=kw, out=prepare_logits_processor(**kw)(None, x)))
x3 = torch.randn(5, 4, 200, generator=g)
obs = {}
for name, kw in (("top_p_3d", dict(temperature=0.7, top_p=0.9)), ("top_k_3d", dict(temperature=0.7, top_k=20))):
    try:
        prepare_logits_processor(**kw)(None, x3)
        obs[name] = "ok"
    except Exception as e:           # noqa: BLE001
        obs[name] = f"{type(e).__name__}: {e}"
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "warpers_sampling.pt")
torch.save(dict(x=x, cases=cases, observations=obs, transformers=__import__("transformers").__version__), path)
print("wrote", path, obs)
