"""ctypes binding of libvispec_b200.so (C ABI in include/vispec_b200.h).

The product path has no CPU fallback: if the library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

c_i32, c_i64, c_f32, c_vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p

VS_TOK_BLOCK = 32
VS_RETRIEVE_LD = 8
VS_MAX_ACCEPT = 8
(ST_CUR_LEN, ST_PREV_LEN, ST_ACCEPT, ST_BEST, ST_NEXT_TOKEN, ST_D_LEN, ST_D_POS, ST_M, ST_N_LEAVES, ST_MAX_DEPTH,
 ST_LAST_ROW, ST_CYCLE, ST_TEMP_BITS, ST_RNG_SEED, ST_RNG_CTR, ST_ROPE_DELTA) = range(16)
ST_SLOTS = 16
EPI_STORE, EPI_RESID, EPI_SWIGLU, EPI_HEADS = range(4)
(PH_NOP, PH_GEMM, PH_EPI, PH_NORM, PH_ROWS, PH_ATTN_S, PH_ATTN_PV, PH_ARGMAX, PH_LSM_STATS, PH_TREE, PH_SPLICE,
 PH_LSM_CAND, PH_ATTN_FUSED, PH_GEMM2, PH_TOPK_HIST) = range(15)
PHASE_NAMES = ("nop", "gemm", "epilogue", "rmsnorm", "rows", "attn_scores", "attn_pv", "argmax", "lsm_stats", "tree",
               "kv_splice", "lsm_candidates", "attn_fused", "slab_gemm", "topk_hist")


class LinearArgs(C.Structure):
    _fields_ = [("X", c_vp), ("ldx", c_i64), ("W", c_vp), ("ldw", c_i64), ("bias", c_vp),
                ("T", c_i32), ("N", c_i32), ("K", c_i32), ("splits", c_i32),
                ("partials", c_vp), ("counters", c_vp), ("epi", c_i32),
                ("Y", c_vp), ("ldy", c_i64), ("R", c_vp), ("ldr", c_i64),
                ("n_q_tiles", c_i32), ("n_k_tiles", c_i32), ("rope", c_i32),
                ("cos", c_vp), ("sin", c_vp), ("pos_ids", c_vp),
                ("Qout", c_vp), ("q_head_stride", c_i64),
                ("Kout", c_vp), ("Vout", c_vp), ("kv_head_stride", c_i64),
                ("kv_row_base", c_vp), ("kv_row_off", c_i32)]


class AttnArgs(C.Structure):
    _fields_ = [("Q", c_vp), ("q_head_stride", c_i64), ("K", c_vp), ("V", c_vp), ("kv_head_stride", c_i64),
                ("n_heads", c_i32), ("n_kv_heads", c_i32), ("Tq", c_i32),
                ("n_ctx", c_vp), ("n_ctx_host", c_i32), ("n_new", c_i32), ("causal", c_i32),
                ("bits", c_vp), ("eager", c_i32), ("scale", c_f32),
                ("scores", c_vp), ("score_ld", c_i64), ("stats", c_vp), ("max_chunks", c_i32),
                ("opart", c_vp), ("pv_splits", c_i32), ("counters", c_vp), ("O", c_vp), ("ldo", c_i64),
                ("prefix_stable", c_i32)]


class TreeBufs(C.Structure):
    _fields_ = [("top_k", c_i32), ("depth", c_i32), ("total", c_i32),
                ("scores_flat", c_vp), ("tokens_flat", c_vp), ("parents", c_vp),
                ("frontier_scores", c_vp), ("frontier_ids", c_vp), ("cs_index", c_vp), ("frontier_bits", c_vp),
                ("top_val", c_vp), ("top_idx", c_vp), ("gather_idx", c_vp), ("draft_pos_ids", c_vp),
                ("draft_tokens", c_vp), ("tree_bits", c_vp), ("tree_pos", c_vp), ("verify_pos", c_vp),
                ("retrieve", c_vp), ("sel", c_vp), ("stable_ids", c_vp), ("stable_pos", c_vp)]


class Dims(C.Structure):
    _fields_ = [("hidden", c_i32), ("inter", c_i32), ("n_heads", c_i32), ("n_kv_heads", c_i32), ("vocab", c_i32),
                ("n_layers", c_i32), ("eps", c_f32)]


class Layer(C.Structure):
    _fields_ = [("ln_in", c_vp), ("w_qkv", c_vp), ("b_qkv", c_vp), ("w_o", c_vp), ("ln_post", c_vp),
                ("w_gateup", c_vp), ("w_down", c_vp), ("k_rows", c_vp), ("v_rows", c_vp), ("kv_head_stride", c_i64)]


class Scratch(C.Structure):
    _fields_ = [("max_rows", c_i32), ("h", c_vp), ("x", c_vp), ("q", c_vp), ("attn", c_vp), ("act", c_vp),
                ("partials", c_vp), ("partials_floats", c_i64), ("counters", c_vp),
                ("scores", c_vp), ("score_ld", c_i64), ("stats", c_vp), ("max_chunks", c_i32),
                ("opart", c_vp), ("pv_splits", c_i32), ("attn_counters", c_vp)]


class Rope(C.Structure):
    _fields_ = [("cos", c_vp), ("sin", c_vp)]


class Target(C.Structure):
    _fields_ = [("dims", Dims), ("layers", C.POINTER(Layer)), ("embed", c_vp), ("norm", c_vp), ("lm_head", c_vp),
                ("rope", Rope), ("sdpa_attention", c_i32)]


class Draft(C.Structure):
    _fields_ = [("dims", Dims), ("layer", Layer), ("embed", c_vp), ("w_fc", c_vp), ("b_fc", c_vp),
                ("w_img_fc", c_vp), ("b_img_fc", c_vp), ("lm_head", c_vp), ("rope", Rope),
                ("x1", c_vp), ("x2", c_vp), ("logits", c_vp)]


class Adaptor(C.Structure):
    _fields_ = [("w_kv", c_vp), ("b_kv", c_vp), ("w_o", c_vp), ("q", c_vp), ("num_q", c_i32)]


class PrefillWs(C.Structure):
    _fields_ = [("max_rows", c_i32), ("img_k", c_vp), ("img_v", c_vp), ("x1", c_vp), ("x2", c_vp), ("fused", c_vp),
                ("adapted", c_vp), ("attn_out", c_vp), ("pos", c_vp), ("packed", c_vp)]


_SIGS = {
    "vs_last_error": (C.c_char_p, []),
    "vs_abi_version": (C.c_int, []),
    "vs_linear_splits": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "vs_linear_partials_floats": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "vs_linear": (C.c_int, [C.POINTER(LinearArgs), c_vp]),
    "vs_rmsnorm": (C.c_int, [c_vp, c_i64, c_vp, c_vp, c_i64, C.c_int, C.c_int, c_f32, c_vp]),
    "vs_gather_rows": (C.c_int, [c_vp, c_i64, c_vp, c_vp, c_i64, C.c_int, C.c_int, c_vp]),
    "vs_embed_rows": (C.c_int, [c_vp, c_i64, c_vp, c_vp, c_i64, C.c_int, C.c_int, C.c_int, c_vp]),
    "vs_broadcast_row": (C.c_int, [c_vp, c_vp, c_i64, C.c_int, C.c_int, c_vp]),
    "vs_argmax_rows": (C.c_int, [c_vp, c_i64, C.c_int, C.c_int, c_vp, c_vp]),
    "vs_logsoftmax_topk": (C.c_int, [c_vp, c_i64, C.c_int, C.c_int, C.c_int, c_vp, c_vp, c_vp, c_vp]),
    "vs_attention": (C.c_int, [C.POINTER(AttnArgs), c_vp]),
    "vs_draft_level0": (C.c_int, [C.POINTER(TreeBufs), c_vp, c_vp]),
    "vs_draft_expand": (C.c_int, [C.POINTER(TreeBufs), c_vp, C.c_int, c_vp]),
    "vs_tree_build": (C.c_int, [C.POINTER(TreeBufs), c_vp, c_vp, C.c_int, c_vp]),
    "vs_accept_greedy": (C.c_int, [C.POINTER(TreeBufs), c_vp, c_vp, c_vp, c_vp]),
    "vs_kv_splice": (C.c_int, [c_vp, C.c_int, C.c_int, c_i64, c_vp, c_vp, C.c_int, C.c_int, c_vp, c_vp]),
    "vs_set_state": (C.c_int, [c_vp, C.c_int, C.c_int, c_vp]),
    "vs_decoder_layer": (C.c_int, [C.POINTER(Dims), C.POINTER(Layer), C.POINTER(Scratch), C.POINTER(Rope), C.c_int,
                                   c_vp, c_vp, C.c_int, c_vp, C.c_int, C.c_int, C.c_int, c_vp, C.c_int, c_vp]),
    "vs_verify_step": (C.c_int, [C.POINTER(Target), C.POINTER(Scratch), c_vp, c_vp, c_vp, C.c_int, c_vp, c_vp, c_vp,
                                 c_vp, c_vp]),
    "vs_draft_forward": (C.c_int, [C.POINTER(Draft), C.POINTER(Scratch), C.c_int, c_vp, c_vp, c_vp, C.c_int, c_vp,
                                   C.c_int, C.c_int, c_vp, c_vp]),
    "vs_img_adaptor_prefill": (C.c_int, [C.POINTER(Draft), C.POINTER(Adaptor), C.POINTER(Scratch), C.POINTER(PrefillWs), c_vp,
                                         c_vp, C.c_int, C.c_int, C.c_int, c_vp, c_vp, c_vp, c_vp]),
    "vs_draft_levels": (C.c_int, [C.POINTER(Draft), C.POINTER(Scratch), C.POINTER(TreeBufs), c_vp, C.c_int, c_vp,
                                  C.c_int, c_vp]),
    "vs_draft_tree": (C.c_int, [C.POINTER(Draft), C.POINTER(Scratch), C.POINTER(TreeBufs), c_vp, c_vp]),
    "vs_spec_cycle": (C.c_int, [C.POINTER(Target), C.POINTER(Scratch), C.POINTER(Draft), C.POINTER(Scratch),
                                C.POINTER(TreeBufs), c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    # ---- programs (persistent decode kernel) ----
    "vs_program_create": (C.c_int, [C.POINTER(c_vp)]),
    "vs_program_destroy": (None, [c_vp]),
    "vs_program_last_error": (C.c_char_p, []),
    "vs_program_set_option": (C.c_int, [c_vp, C.c_char_p, C.c_int]),
    "vs_program_num_phases": (C.c_int, [c_vp]),
    "vs_program_phase_types": (C.c_int, [c_vp, c_vp, C.c_int]),
    "vs_program_workspace_bytes": (C.c_size_t, [c_vp]),
    "vs_program_commit": (C.c_int, [c_vp, c_vp, C.c_size_t, c_vp]),
    "vs_program_run": (C.c_int, [c_vp, c_vp]),
    "vs_program_read_trace": (C.c_int, [c_vp, c_vp, C.c_int]),
    "vs_program_read_trace_cta": (C.c_int, [c_vp, c_vp, C.c_int]),
    "vs_program_register_tiled": (C.c_int, [c_vp, c_vp, c_vp]),
    "vs_device_sms": (C.c_int, []),
    "vs_slab_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "vs_slab_plan": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, c_vp, C.c_int]),
    "vs_slab_pack": (C.c_int, [c_vp, c_i64, C.c_int, C.c_int, C.c_int, C.c_int, c_vp, c_vp]),
    "vs_program_register_slab": (C.c_int, [c_vp, c_vp, c_vp, C.c_int, C.c_int]),
    "vs_attention_fused_capacity": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "vs_program_add_linear": (C.c_int, [c_vp, C.POINTER(LinearArgs)]),
    "vs_program_add_rmsnorm": (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, c_i64, C.c_int, C.c_int, c_f32]),
    "vs_program_add_gather_rows": (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, c_i64, C.c_int, C.c_int]),
    "vs_program_add_embed_rows": (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, c_i64, C.c_int, C.c_int, C.c_int]),
    "vs_program_add_broadcast_row": (C.c_int, [c_vp, c_vp, c_vp, c_i64, C.c_int, C.c_int]),
    "vs_program_add_argmax_rows": (C.c_int, [c_vp, c_vp, c_i64, C.c_int, C.c_int, c_vp]),
    "vs_program_add_logsoftmax_topk": (C.c_int, [c_vp, c_vp, c_i64, C.c_int, C.c_int, C.c_int, c_vp, c_vp, c_vp]),
    "vs_program_add_attention": (C.c_int, [c_vp, C.POINTER(AttnArgs)]),
    "vs_program_add_draft_level0": (C.c_int, [c_vp, C.POINTER(TreeBufs), c_vp]),
    "vs_program_add_draft_expand": (C.c_int, [c_vp, C.POINTER(TreeBufs), c_vp, C.c_int]),
    "vs_program_add_tree_build": (C.c_int, [c_vp, C.POINTER(TreeBufs), c_vp, c_vp, C.c_int]),
    "vs_program_add_accept_greedy": (C.c_int, [c_vp, C.POINTER(TreeBufs), c_vp, c_vp, c_vp]),
    "vs_program_add_kv_splice": (C.c_int, [c_vp, c_vp, C.c_int, C.c_int, c_i64, c_vp, c_vp]),
    "vs_program_add_accept_sampling": (C.c_int, [c_vp, C.POINTER(TreeBufs), c_vp, c_vp, c_i64, C.c_int, c_f32, c_vp, c_vp]),
    "vs_program_add_spec_cycle_sampling": (C.c_int, [c_vp, C.POINTER(Target), C.POINTER(Scratch), C.POINTER(Draft),
                                                     C.POINTER(Scratch), C.POINTER(TreeBufs), c_vp, c_vp, c_i64, c_vp,
                                                     c_vp, c_f32, c_vp]),
    "vs_program_add_decoder_layer": (C.c_int, [c_vp, C.POINTER(Dims), C.POINTER(Layer), C.POINTER(Scratch), C.POINTER(Rope),
                                               C.c_int, c_vp, c_vp, C.c_int, c_vp, C.c_int, C.c_int, C.c_int, c_vp, C.c_int]),
    "vs_program_add_verify_step": (C.c_int, [c_vp, C.POINTER(Target), C.POINTER(Scratch), c_vp, c_vp, c_vp, C.c_int, c_vp,
                                             c_vp, c_vp, c_vp]),
    "vs_program_add_draft_forward": (C.c_int, [c_vp, C.POINTER(Draft), C.POINTER(Scratch), C.c_int, c_vp, c_vp, c_vp,
                                               C.c_int, c_vp, C.c_int, C.c_int, c_vp]),
    "vs_program_add_draft_levels": (C.c_int, [c_vp, C.POINTER(Draft), C.POINTER(Scratch), C.POINTER(TreeBufs), c_vp,
                                              C.c_int, c_vp, C.c_int]),
    "vs_program_add_draft_tree": (C.c_int, [c_vp, C.POINTER(Draft), C.POINTER(Scratch), C.POINTER(TreeBufs), c_vp]),
    "vs_program_add_spec_cycle": (C.c_int, [c_vp, C.POINTER(Target), C.POINTER(Scratch), C.POINTER(Draft),
                                            C.POINTER(Scratch), C.POINTER(TreeBufs), c_vp, c_vp, c_i64, c_vp, c_vp, c_vp,
                                            c_vp]),
    "vs_launch_count": (C.c_longlong, [C.c_int]),
    "vs_profile_start": (C.c_int, [C.c_int]),
    "vs_profile_stop": (C.c_int, [c_vp, c_vp, c_vp, C.c_int]),
}
EXPORTS = tuple(_SIGS)

_lib = None


def lib_path() -> str:
    # VS_LIB_PATH: an alternative build of the same library (kernel A/B experiments); still no fallback of any kind
    return os.environ.get("VS_LIB_PATH") or _build.LIB_PATH


def load():
    """Load the shared library (never falls back to anything else)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(vispec_b200 has no CPU or PyTorch fallback)")
    lib = C.CDLL(path)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.vs_abi_version() != 2:
        raise RuntimeError("libvispec_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().vs_last_error().decode(errors="replace")
        raise RuntimeError(f"vispec_b200 {what} failed (code {rc}): {msg}")


def pcheck(rc: int, what: str = "") -> None:
    """Like check() for the vs_program_* entry points (their messages come from vs_program_last_error)."""
    if rc != 0:
        msg = load().vs_program_last_error().decode(errors="replace")
        raise RuntimeError(f"vispec_b200 {what} failed (code {rc}): {msg}")


def ptr(t) -> int:
    """Device (or host) address of a torch tensor, None -> NULL."""
    return None if t is None else t.data_ptr()
