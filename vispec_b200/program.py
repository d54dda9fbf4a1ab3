"""Python handle of a vs_program (include/vispec_b200.h, "Programs"): a list of phases that ONE launch of the persistent
decode kernel executes.  PyTorch only provides the workspace allocation and the stream."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch

from . import _lib as L


class Program:
    """Build with the `add_*` methods (thin wrappers of `vs_program_add_*`; `p.lib` exposes the rest), `commit()`, then
    `run()` any number of times."""

    def __init__(self, device, pf_ahead: Optional[int] = None, trace: bool = False, attn_fused: bool = False, **options):
        if torch.device(device).type != "cuda":
            raise RuntimeError("vispec_b200 runs on CUDA devices only (no CPU fallback)")
        self.lib = L.load()
        self.device = torch.device(device)
        h = C.c_void_p()
        L.pcheck(self.lib.vs_program_create(C.byref(h)), "vs_program_create")
        self.h = h
        self.ws: Optional[torch.Tensor] = None
        self.trace = trace
        if pf_ahead is not None:
            self.set_option("pf_ahead", pf_ahead)
        if trace:
            self.set_option("trace", 1)
        if attn_fused:
            self.set_option("attn_fused", 1)
        for name, value in options.items():        # any other vs_program_set_option switch (kv_two_loads, attn_prefetch, ...)
            self.set_option(name, value)

    def __del__(self):
        h, self.h = getattr(self, "h", None), None
        if h:
            self.lib.vs_program_destroy(h)

    def set_option(self, name: str, value: int):
        L.pcheck(self.lib.vs_program_set_option(self.h, name.encode(), int(value)), "vs_program_set_option")

    def register_tiled(self, pairs):
        """Streamed copies of decode weights.  Items are (row-major weight, copy[, epi]): epi >= 0 -> `copy` is the slab
        image of engine.pack_weight_slab for that epilogue class (whole rows per CTA, epilogue from TMEM); epi < 0 or
        absent -> the tiled copy of engine.pack_weight_tiles (16 KB tiles, stream-K).  Linears added afterwards that use
        one of these weights stream the copy instead of the row-major matrix."""
        self._tiled_refs = getattr(self, "_tiled_refs", []) + list(pairs)
        with torch.cuda.device(self.device):
            for item in pairs:
                w, wt = item[0], item[1]
                epi = item[2] if len(item) > 2 else -1
                if epi >= 0:
                    L.pcheck(self.lib.vs_program_register_slab(self.h, w.data_ptr(), wt.data_ptr(), epi, 0),
                             "vs_program_register_slab")
                else:
                    L.pcheck(self.lib.vs_program_register_tiled(self.h, w.data_ptr(), wt.data_ptr()),
                             "vs_program_register_tiled")
        return self

    def add(self, fn_name: str, *args):
        """Call `vs_program_add_<fn_name>(program, *args)`."""
        fn = getattr(self.lib, "vs_program_add_" + fn_name)
        with torch.cuda.device(self.device):          # the grid size is the SM count of the program's device
            L.pcheck(fn(self.h, *args), "vs_program_add_" + fn_name)
        return self

    @property
    def n_phases(self) -> int:
        return self.lib.vs_program_num_phases(self.h)

    def phase_types(self) -> List[int]:
        n = self.n_phases
        buf = (C.c_int32 * n)()
        self.lib.vs_program_phase_types(self.h, buf, n)
        return list(buf)

    def commit(self):
        nbytes = self.lib.vs_program_workspace_bytes(self.h)
        self.ws = torch.zeros(nbytes + 256, dtype=torch.uint8, device=self.device)
        off = (-self.ws.data_ptr()) % 256
        self._ws_ptr = self.ws.data_ptr() + off
        with torch.cuda.device(self.device):
            L.pcheck(self.lib.vs_program_commit(self.h, self._ws_ptr, nbytes, torch.cuda.current_stream().cuda_stream),
                     "vs_program_commit")
        return self

    def run(self, stream: Optional[int] = None):
        if stream is None:
            stream = torch.cuda.current_stream(self.device).cuda_stream
        if torch.cuda.current_device() != self.device.index and self.device.index is not None:
            with torch.cuda.device(self.device):
                L.pcheck(self.lib.vs_program_run(self.h, stream), "vs_program_run")
            return
        L.pcheck(self.lib.vs_program_run(self.h, stream), "vs_program_run")

    def read_trace_us(self) -> List[float]:
        """Per-phase durations (µs) of the last run as seen by CTA 0: phase i = stamp[i] - stamp[i-1]; the first entry is
        measured from the end of phase 0 and reported as 0."""
        n = self.n_phases
        buf = (C.c_ulonglong * n)()
        got = self.lib.vs_program_read_trace(self.h, buf, n)
        if got < 0:
            L.pcheck(got, "vs_program_read_trace")
        st = list(buf)[:got]
        return [0.0] + [(st[i] - st[i - 1]) / 1e3 for i in range(1, len(st))]

    def read_trace_detail(self):
        """(end[p], done[p][cta]) in ns: `end` = barrier completion of phase p seen by CTA 0, `done` = when each CTA had
        finished its own work items of phase p (before the barrier)."""
        n = self.n_phases
        buf = (C.c_ulonglong * n)()
        L.pcheck(min(0, self.lib.vs_program_read_trace(self.h, buf, n)), "vs_program_read_trace")
        cap = n * 4096
        big = (C.c_ulonglong * cap)()
        g = self.lib.vs_program_read_trace_cta(self.h, big, cap)
        if g < 0:
            L.pcheck(g, "vs_program_read_trace_cta")
        done = [[big[p * g + b] for b in range(g)] for p in range(n)]
        return list(buf), done
