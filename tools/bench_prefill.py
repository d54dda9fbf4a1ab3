"""Target prompt pass at BASELINE config-2 shapes (704 rows, LLaVA-7B, 32 layers): library GEMMs (cuBLAS, default) vs the
library's own kernels (VS_PREFILL=native), CUDA events."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_llava7b  # noqa: E402

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 32
model, ids, emb, mask = build_llava7b(device="cuda", seed=0, layers=layers)
eng = model.base_model_engine


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print(f"rows={emb.shape[1]} layers={layers} cublas_ms={timed(lambda: eng.prefill_library(emb)):.2f} "
      f"native_ms={timed(lambda: eng.prefill_native(emb)):.2f}")
