"""Independent Python restatement of the slab weight layout (vispec_b200/csrc/slab.h) and of how the program kernel reads
it: used by the CPU test (plan + index consistency: pack -> stage copy -> address-swizzled MMA read -> accumulator lane ->
output feature) and by the GPU test that compares `vs_slab_pack` with it byte for byte."""
from __future__ import annotations

import ctypes as C

import numpy as np

FIELDS = ("pair", "packed", "E", "n_tiles", "passes", "rows_alloc", "blk_bytes", "blk_smem", "nk", "kc", "x_off", "e_per",
          "e_rem", "grid")


def plan(lib, N, K, epi, grid):
    buf = (C.c_int32 * 14)()
    n = lib.vs_slab_plan(N, K, epi, grid, buf, 14)
    if n == 0:
        return None
    return dict(zip(FIELDS, list(buf)))


def unit_range(p, u):
    lo = u * p["e_per"] + min(u, p["e_rem"])
    return lo, p["e_per"] + (1 if u < p["e_rem"] else 0)


def source_row(p, u, t, r):
    """Row of W held by lane r of accumulator t of unit u, or -1."""
    lo, cnt = unit_range(p, u)
    if p["pair"]:
        if r >= cnt:
            return -1
        e = lo + r
        return (e // 64) * 128 + e % 64 + 64 * t
    e = lo + t * p["rows_alloc"] + r
    return e if (r < p["rows_alloc"] and e < lo + cnt) else -1


def pack(p, W: np.ndarray) -> np.ndarray:
    """W [N, K] uint16 (bf16 bits) -> the slab image as uint16 (little endian bf16), following slab.h."""
    N, K = W.shape
    units = p["grid"] * p["passes"]
    unit_stride = p["nk"] * p["n_tiles"] * p["blk_bytes"]
    out = np.zeros(units * unit_stride // 2, dtype=np.uint16)
    for u in range(units):
        for j in range(p["nk"]):
            jj = j % p["kc"]
            for t in range(p["n_tiles"]):
                base = u * unit_stride + (j * p["n_tiles"] + t) * p["blk_bytes"]
                for r in range(p["rows_alloc"]):
                    n = source_row(p, u, t, r)
                    if n < 0 or n >= N:
                        continue
                    o = (jj * p["n_tiles"] + t) * p["blk_smem"] + r * 128     # offset of the row inside its stage
                    for c in range(8):
                        cs = c ^ ((r & 7) if p["packed"] == 2 else ((o >> 7) & 7))
                        dst = (base + r * 128 + cs * 16) // 2
                        out[dst:dst + 8] = W[n, j * 64 + c * 8:j * 64 + c * 8 + 8]
    return out


def emulate_gemm(p, slab: np.ndarray, X: np.ndarray, N: int) -> np.ndarray:
    """What the kernel computes from the image: per CTA / pass, stage copies into a 32 KB slot, MMA reads with the
    address-based 128-byte swizzle (row i of a block that starts at byte `a` of the slot sits at a + i * 128, its 16-byte
    chunk c at chunk position c ^ (address bits [7, 10))), fp32 accumulation, then the lane -> feature map of
    `slab_epilogue`.  X [T, K] float32; returns Y [T, N] float32 (row pairs are written to their own rows)."""
    T, K = X.shape
    Y = np.zeros((T, N), dtype=np.float32)
    seen = np.zeros(N, dtype=np.int32)
    unit_stride = p["nk"] * p["n_tiles"] * p["blk_bytes"] // 2
    slot = np.zeros(32768 // 2, dtype=np.uint16)
    for b in range(p["grid"]):
        for ps in range(p["passes"]):
            u = b * p["passes"] + ps
            lo, cnt = unit_range(p, u)
            if cnt == 0:
                continue
            acc = np.zeros((p["n_tiles"], 128, T), dtype=np.float32)
            src = u * unit_stride
            for j0 in range(0, p["nk"], p["kc"]):
                nj = min(p["kc"], p["nk"] - j0)
                nblk = nj * p["n_tiles"]
                slot[:] = 0xFFFF                                    # stale bytes: NaN patterns, must never reach a valid lane
                for i in range(nblk):                               # the bulk copies of the stage
                    d = i * p["blk_smem"] // 2
                    slot[d:d + p["blk_bytes"] // 2] = slab[src + i * p["blk_bytes"] // 2: src + (i + 1) * p["blk_bytes"] // 2]
                src += nblk * p["blk_bytes"] // 2
                for jj in range(nj):
                    xk = X[:, (j0 + jj) * 64:(j0 + jj + 1) * 64]
                    for t in range(p["n_tiles"]):
                        a = (jj * p["n_tiles"] + t) * p["blk_smem"]
                        for r in range(p["rows_alloc"]):            # (lanes beyond hold garbage and are never read)
                            row = np.empty(64, dtype=np.uint16)
                            addr = a + r * 128
                            for c in range(8):     # packed == 2: pattern relative to the block (descriptor base offset)
                                phase = (r & 7) if p["packed"] == 2 else ((addr >> 7) & 7)
                                phys = addr + ((c ^ phase) << 4)
                                row[c * 8:c * 8 + 8] = slot[phys // 2: phys // 2 + 8]
                            wf = (row.astype(np.uint32) << 16).view(np.float32)
                            acc[t, r] += xk @ wf
            for t in range(p["n_tiles"]):
                for r in range(128):
                    n = source_row(p, u, t, r) if r < p["rows_alloc"] else -1
                    if 0 <= n < N:
                        Y[:, n] = acc[t, r]
                        seen[n] += 1
    assert (seen == 1).all(), "every output row must be owned by exactly one (CTA, pass, accumulator, lane)"
    return Y
