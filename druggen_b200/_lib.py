"""ctypes binding of libdruggen_b200.so (the C-ABI declared in include/druggen_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` (``make -C druggen_b200/csrc``).
There is no fallback: if the library is missing or a launch is rejected this raises.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdruggen_b200.so")
PREC = {"fp32": 0, "bf16": 1, "bf16x3": 2}

_P, _LL, _I, _F = C.c_void_p, C.c_longlong, C.c_int, C.c_float

# name -> argument ctypes (every entry returns int); mirrors include/druggen_b200.h
SIGNATURES = {
    "dg_rows_gemm": [_P, _P, _I, _P, _I, _P, _P, _P, _LL, _I, _I, _I, _I, _P],
    "dg_gemm_tn": [_P, _P, _P, _P, _LL, _I, _I, _I, _I, _P],
    "dg_colsum": [_P, _P, _LL, _I, _P],
    "dg_gate_mul": [_P, _P, _P, _LL, _P],
    "dg_add_ln_fwd": [_P, _P, _P, _P, _P, _LL, _I, _F, _P],
    "dg_add_ln_bwd": [_P, _P, _P, _P, _P, _P, _P, _LL, _I, _F, _I, _P],
    "dg_add_ln_bwd_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _LL, _I, _F, _P],
    "dg_modulate_fwd": [_P, _P, _P, _F, _P, _I, _I, _I, _P],
    "dg_modulate_bwd": [_P, _P, _P, _P, _F, _P, _P, _P, _I, _I, _I, _P],
    "dg_modulate_bwd_bwd": [_P, _P, _P, _P, _P, _P, _P, _F, _P, _P, _P, _P, _I, _I, _I, _P],
    "dg_softmax_agg_fwd": [_P, _P, _P, _I, _I, _I, _P],
    "dg_softmax_agg_bwd": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "dg_softmax_agg_bwd_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    "dg_attn_scores_fwd": [_P, _P, _P, _P, _F, _P, _P, _P, _P, _I, _I, _I, _P],
    "dg_attn_scores_bwd": [_P, _P, _P, _P, _P, _P, _F, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "dg_softmax_agg16_fwd": [_P, _P, _P, _P, _P, _I, _I, _I, _P],
    "dg_attn_edge_fwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _F, _P, _P, _P, _P, _I, _I, _I, _F, _P, _LL, _P],
    "dg_mlp_bwd_ln": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _LL, _I, _I, _F, _P, _LL, _P],
    "dg_mlp_bwd_dgrad": [_P, _P, _P, _P, _P, _P, _P, _LL, _I, _I, _P, _LL, _P],
    "dg_mlp_fwd": [_P, _P, _P, _P, _P, _P, _P, _P, _LL, _I, _I, _F, _P, _LL, _P],
    "dg_symmetrize": [_P, _P, _I, _I, _I, _P],
    "dg_label2onehot": [_P, _I, _P, _LL, _I, _P],
    "dg_argmax_last": [_P, _P, _LL, _I, _P],
    "dg_embed_labels_fwd": [_P, _I, _P, _P, _LL, _I, _I, _I, _I, _P],
    "dg_embed_labels_bwd": [_P, _I, _P, _P, _LL, _I, _I, _I, _I, _P],
    "dg_gp_interp": [_P, _I, _P, _P, _P, _LL, _LL, _I, _P],
    "dg_gp_penalty": [_P, _P, _P, _P, _P, _I, _LL, _LL, _P],
    "dg_gp_penalty_bwd": [_P, _P, _P, _P, _I, _LL, _P],
    "dg_readout_argmax": [_P, _P, _P, _P, _P, _I, _LL, _I, _I, _P],
    "dg_adamw_flat": [_P, _P, _P, _P, _P, _I, _F, _F, _F, _F, _F, _P],
    "dg_to_dense_adj": [_P, _P, _P, _P, _P, _LL, _LL, _I, _I, _P],
    "dg_narrow_labels": [_P, _P, _LL, _I, _P],
    "dg_pack_bits": [_P, _I, _P, _P, _LL, _I, _P],
    "dg_tanimoto_agg": [_P, _P, _LL, _P, _P, _LL, _I, _I, _F, _P, _P, _P],
    # block-level entry points: (io table, params table[, grads table], B, N, D, H, heads, flags, eps, workspace, bytes)
    "dg_block_fwd": [_P, _P, _I, _I, _I, _I, _I, _I, _F, _P, _LL, _P],
    "dg_block_bwd": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _P, _LL, _P],
    "dg_block_bwd_bwd": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _P, _LL, _P],
    "dg_encoder_fwd": [_P, _P, _P, _P, _P, _I, _P, _I, _I, _I, _I, _I, _I, _F, _P, _LL, _P],
}
INFO_SYMBOLS = ("dg_abi_version", "dg_last_error", "dg_has_tcgen05", "dg_set_option", "dg_get_option", "dg_debug_chain_profile",
                "dg_label_error", "dg_native_launches", "dg_probe_set", "dg_probe_read", "dg_debug_trace", "dg_debug_trace_read")
# slots of the block-level entry points' buffer table, in the order of the DG_BLK_* enum of include/druggen_b200.h
BLK_SLOTS = ("X", "Y", "X_OUT", "Y_OUT", "X1", "Q", "K", "V", "G", "ON", "X3", "STAT_M", "STAT_INV", "Y3", "A16", "E", "Z4",
             "DXO", "DYO", "DX", "DY", "N_DZ", "N_DX3", "N_DZ3", "N_DG", "N_DQ", "N_DK", "N_DV", "N_T0", "N_T1", "N_H", "N_MASK",
             "E_A", "E_B", "E_H", "E_MASK", "SCRATCH", "UX", "UY", "C_X", "C_Y", "C_DXO", "C_DYO", "N_ARENA", "N_H2", "N_H3",
             "ES0", "ES1", "ES2", "ES3", "ES4", "ES5", "ES6", "ES7", "ES8", "E_H2", "E_H3", "WT")
BB_NODE_SLOTS = 28
BLK = {name: i for i, name in enumerate(BLK_SLOTS)}
BLKF_EDGE_OUT, BLKF_KEEP, BLKF_STATS = 1, 2, 4
BLOCK_PARAMS = 30
ABI_VERSION = 6
OPT_L2_PREFETCH = 0
OPT_ATTN_BWD = 1                                      # 0 = TMA-fed ring kernel where it applies (default), 1 = 4-warp kernel
PF_ALL, PF_DEFAULT, PF_CHAIN_KEEP = 63, 12, 64        # DG_PF_* bit masks (include/druggen_b200.h)

_lib = None
_backend = None


def load():
    """dlopen the library and type every symbol (no device work)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `make -C druggen_b200/csrc` "
                "(or __graft_entry__.build()); druggen_b200 has no CPU / eager fallback")
        lib = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes, fn.restype = args, _I
        lib.dg_abi_version.restype = _I
        lib.dg_has_tcgen05.restype = _I
        lib.dg_last_error.restype = C.c_char_p
        lib.dg_set_option.argtypes, lib.dg_set_option.restype = [_I, _I], _I
        lib.dg_get_option.argtypes, lib.dg_get_option.restype = [_I], _I
        lib.dg_debug_chain_profile.argtypes, lib.dg_debug_chain_profile.restype = [_P], _I
        lib.dg_label_error.argtypes, lib.dg_label_error.restype = [_I], _I
        lib.dg_native_launches.argtypes, lib.dg_native_launches.restype = [], _LL
        lib.dg_probe_set.argtypes, lib.dg_probe_set.restype = [C.c_char_p], _I
        lib.dg_probe_read.argtypes, lib.dg_probe_read.restype = [C.POINTER(_LL), C.POINTER(C.c_double)], _I
        lib.dg_debug_trace.argtypes, lib.dg_debug_trace.restype = [_I], _I
        lib.dg_debug_trace_read.argtypes, lib.dg_debug_trace_read.restype = [C.c_char_p, _LL], _LL
        if lib.dg_abi_version() != ABI_VERSION:
            raise RuntimeError("libdruggen_b200.so ABI version mismatch")
        if os.environ.get("DRUGGEN_B200_L2_PREFETCH") is not None:       # tuning switch, default on
            lib.dg_set_option(OPT_L2_PREFETCH, int(os.environ["DRUGGEN_B200_L2_PREFETCH"]))
        if os.environ.get("DRUGGEN_B200_ATTN_BWD") is not None:
            lib.dg_set_option(OPT_ATTN_BWD, int(os.environ["DRUGGEN_B200_ATTN_BWD"]))
        _lib = lib
    return _lib


def _ptr(t):
    return None if t is None else t.data_ptr()


def _nbytes(*ts):
    return sum(t.numel() * t.element_size() for t in ts if t is not None)


class CudaBackend:
    """Launch table used by kernels.py: torch tensors -> raw pointers + current stream.

    Optional per-kernel timing (bench.py): with ``profile_all`` or ``profile_only == key`` every
    matching launch is bracketed by CUDA events on the launching stream; ``profile_summary()``
    returns, per key, launches / total ms / algorithmic bytes and FLOPs."""

    def __init__(self, lib):
        self.lib = lib
        self._py_launches = 0      # kernel launches issued one by one through this table
        self._native0 = lib.dg_native_launches()
        self.device = None         # device of the tensors of the launch being issued (set by kernels._chk)
        self.profile_all = False
        self._profile_only = None
        self._prof = {}
        self._per_launch = {}      # key -> (flops, bytes, algorithmic bytes, bound) of ONE launch (for launches the library issues itself)

    @property
    def launches(self):
        """Kernel launches so far (bench.py reports it): those issued from Python plus those the block-level entry points
        (dg_block_fwd / dg_block_bwd / dg_encoder_fwd) issued themselves."""
        return self._py_launches + self.lib.dg_native_launches() - self._native0

    @property
    def profile_only(self):
        return self._profile_only

    @profile_only.setter
    def profile_only(self, key):
        self._profile_only = key
        self.lib.dg_probe_set(key.encode() if key else None)     # the same key times the launches the library issues itself

    def profile_reset(self):
        self._prof = {}
        self.lib.dg_probe_read(None, None)

    def profile_summary(self):
        torch.cuda.synchronize()
        out = {}
        for key, rec in self._prof.items():
            ms = sum(a.elapsed_time(b) for a, b in rec["events"])
            out[key] = {"key": key, "n": len(rec["events"]), "ms": ms, "bytes": rec["bytes"], "alg_bytes": rec["alg_bytes"],
                        "flops": rec["flops"], "bound": rec["bound"]}
        key = self._profile_only
        if key is not None and key in self._per_launch:
            n, ms = _LL(0), C.c_double(0.0)
            self.lib.dg_probe_read(C.byref(n), C.byref(ms))
            if n.value:
                flops, nbytes, alg, bound = self._per_launch[key]
                rec = out.setdefault(key, {"key": key, "n": 0, "ms": 0.0, "bytes": 0, "alg_bytes": 0, "flops": 0, "bound": bound})
                rec["n"] += n.value
                rec["ms"] += ms.value
                rec["bytes"] += nbytes * n.value
                rec["alg_bytes"] += alg * n.value
                rec["flops"] += flops * n.value
        return out

    def native_blocks(self) -> bool:
        """May the block-level entry points run?  Not while every launch is being timed from Python (bench.py's kernel table)."""
        return not self.profile_all

    def _native(self, name, *args):
        dev = self.device
        if dev is not None and dev.index is not None and dev.index != torch.cuda.current_device():
            with torch.cuda.device(dev):
                rc = getattr(self.lib, name)(*args, torch.cuda.current_stream().cuda_stream)
        else:
            rc = getattr(self.lib, name)(*args, torch.cuda.current_stream().cuda_stream)
        if rc != 0:
            raise RuntimeError(f"{name} rejected: {self.lib.dg_last_error().decode()}")

    @staticmethod
    def _table(io):
        tab = (C.c_void_p * len(BLK_SLOTS))()
        for name, t in io.items():
            if t is not None:
                tab[BLK[name]] = t.data_ptr()
        return tab

    @staticmethod
    def _ptrs(ts):
        return (C.c_void_p * len(ts))(*[None if t is None else t.data_ptr() for t in ts])

    def block_fwd(self, io, params, b, n, d, h, heads, flags, eps, ws):
        self._native("dg_block_fwd", self._table(io), self._ptrs(params), b, n, d, h, heads, flags, eps, _ptr(ws), ws.numel())

    def block_bwd(self, io, params, grads, b, n, d, h, heads, flags, eps, ws):
        self._native("dg_block_bwd", self._table(io), self._ptrs(params), None if grads is None else self._ptrs(grads), b, n, d, h,
                     heads, flags, eps, _ptr(ws), ws.numel())

    def block_bwd_bwd(self, io, params, grads, b, n, d, h, heads, flags, eps, ws):
        self._native("dg_block_bwd_bwd", self._table(io), self._ptrs(params), self._ptrs(grads), b, n, d, h, heads, flags, eps,
                     _ptr(ws), ws.numel())

    def encoder_fwd(self, x, y, x_out, y_out, params, depth, scratch, b, n, d, h, heads, last_edge_out, eps, ws):
        self._native("dg_encoder_fwd", _ptr(x), _ptr(y), _ptr(x_out), _ptr(y_out), self._ptrs(params), depth, self._table(scratch),
                     b, n, d, h, heads, int(last_edge_out), eps, _ptr(ws), ws.numel())

    @staticmethod
    def profile_name(rec):
        return rec["key"]

    def _call(self, name, meta, *args):
        self._py_launches += 1
        key, flops, nbytes, bound = meta[:4]
        alg = meta[4] if len(meta) > 4 else nbytes          # SURVEY 8(d) bytes: the kernel's inputs and outputs, no operand spills
        timed = self.profile_all or (self._profile_only is not None and self._profile_only == key)
        if timed:
            self._per_launch[key] = (flops, nbytes, alg, bound)
        dev = self.device
        if dev is not None and dev.index is not None and dev.index != torch.cuda.current_device():
            # tensors on a device that is not the current one (nn.DataParallel replicas, models on cuda:1 without
            # set_device): launch there, on that device's current stream
            with torch.cuda.device(dev):
                return self._launch(name, key, flops, nbytes, alg, bound, timed, args)
        return self._launch(name, key, flops, nbytes, alg, bound, timed, args)

    def _launch(self, name, key, flops, nbytes, alg, bound, timed, args):
        stream = torch.cuda.current_stream()
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
        rc = getattr(self.lib, name)(*args, stream.cuda_stream)
        if rc != 0:
            raise RuntimeError(f"{name} rejected: {self.lib.dg_last_error().decode()}")
        if timed:
            e1.record(stream)
            rec = self._prof.setdefault(key, {"events": [], "bytes": 0, "alg_bytes": 0, "flops": 0, "bound": bound})
            rec["events"].append((e0, e1))
            rec["bytes"] += nbytes
            rec["alg_bytes"] += alg
            rec["flops"] += flops

    def rows_gemm(self, a, w, w_is_nk, bias, relu, gate, out, prec, resid=None):
        r, k = a.shape
        n = out.shape[1]
        tag = ("+gate" if gate is not None else "") + ("+resid" if resid is not None else "") + ("+relu" if relu else "")
        tag += (",a16" if a.dtype == torch.bfloat16 else "") + (",o16" if out.dtype == torch.bfloat16 else "")
        meta = (f"rows_gemm[R={r},K={k},N={n},{prec}{tag}]", 2 * r * k * n, _nbytes(a, out, gate, resid), "hbm")
        bf = torch.bfloat16
        flags = (1 if a.dtype == bf else 0) | (2 if out.dtype == bf else 0) | (4 if gate is not None and gate.dtype == bf else 0)
        self._call("dg_rows_gemm", meta, _ptr(a), _ptr(w), int(w_is_nk), _ptr(bias), int(relu), _ptr(gate), _ptr(resid),
                   _ptr(out), r, k, n, PREC[prec], flags)

    def gemm_tn(self, a, b, out, accumulate, prec, colsum_a=None):
        r, m, n = a.shape[0], a.shape[1], b.shape[1]
        tag = (",a16" if a.dtype == torch.bfloat16 else "") + (",b16" if b.dtype == torch.bfloat16 else "")
        meta = (f"gemm_tn[R={r},M={m},N={n},{prec}{tag}]", 2 * r * m * n, _nbytes(a, b), "hbm")
        flags = (1 if a.dtype == torch.bfloat16 else 0) | (2 if b.dtype == torch.bfloat16 else 0)
        self._call("dg_gemm_tn", meta, _ptr(a), _ptr(b), _ptr(out), _ptr(colsum_a), r, m, n, PREC[prec], flags)

    def colsum(self, a, out):
        self._call("dg_colsum", ("colsum", 0, _nbytes(a), "hbm"), _ptr(a), _ptr(out), a.shape[0], a.shape[1])

    def gate_mul(self, x, ref, out):
        self._call("dg_gate_mul", ("gate_mul", 0, _nbytes(x, ref, out), "hbm"), _ptr(x), _ptr(ref), _ptr(out), x.numel())

    def add_ln_fwd(self, a, b, gamma, beta, out, eps):
        d = a.shape[-1]
        self._call("dg_add_ln_fwd", ("add_ln_fwd", 0, _nbytes(a, b, out), "hbm"), _ptr(a), _ptr(b), _ptr(gamma), _ptr(beta),
                   _ptr(out), a.numel() // d, d, eps)

    def add_ln_bwd(self, dy, a, b, gamma, dz, dgamma, dbeta, eps, accumulate=False):
        d = a.shape[-1]
        self._call("dg_add_ln_bwd", ("add_ln_bwd", 0, _nbytes(dy, a, b, dz) + (_nbytes(dz) if accumulate else 0), "hbm"), _ptr(dy), _ptr(a),
                   _ptr(b), _ptr(gamma), _ptr(dz), _ptr(dgamma), _ptr(dbeta), a.numel() // d, d, eps, int(accumulate))

    def add_ln_bwd_bwd(self, u, vg, vb, dy, a, b, gamma, g_dy, g_z, g_gamma, eps):
        d = a.shape[-1]
        self._call("dg_add_ln_bwd_bwd", ("add_ln_bwd_bwd", 0, _nbytes(u, dy, a, b, g_dy, g_z), "hbm"), _ptr(u), _ptr(vg),
                   _ptr(vb), _ptr(dy), _ptr(a), _ptr(b), _ptr(gamma), _ptr(g_dy), _ptr(g_z), _ptr(g_gamma),
                   a.numel() // d, d, eps)

    def modulate_fwd(self, q, k, e, c, out):
        b, n, d = q.shape
        self._call("dg_modulate_fwd", ("modulate_fwd", 0, _nbytes(e, out), "hbm"), _ptr(q), _ptr(k), _ptr(e), c, _ptr(out), b, n, d)

    def modulate_bwd(self, da, q, k, e, c, dq, dk, de):
        b, n, d = q.shape
        self._call("dg_modulate_bwd", ("modulate_bwd", 0, _nbytes(da, e, de), "hbm"), _ptr(da), _ptr(q), _ptr(k), _ptr(e), c,
                   _ptr(dq), _ptr(dk), _ptr(de), b, n, d)

    def modulate_bwd_bwd(self, uq, uk, ue, da, q, k, e, c, g_da, g_q, g_k, g_e):
        b, n, d = q.shape
        self._call("dg_modulate_bwd_bwd", ("modulate_bwd_bwd", 0, _nbytes(ue, da, e, g_da, g_e), "hbm"), _ptr(uq), _ptr(uk),
                   _ptr(ue), _ptr(da), _ptr(q), _ptr(k), _ptr(e), c, _ptr(g_da), _ptr(g_q), _ptr(g_k), _ptr(g_e), b, n, d)

    def softmax_agg_fwd(self, a, v, out):
        b, n, d = v.shape
        self._call("dg_softmax_agg_fwd", ("softmax_agg_fwd", 0, _nbytes(a), "hbm"), _ptr(a), _ptr(v), _ptr(out), b, n, d)

    def softmax_agg_bwd(self, dg, a, v, da, dv, accumulate=False):
        b, n, d = v.shape
        self._call("dg_softmax_agg_bwd", ("softmax_agg_bwd", 0, _nbytes(a, da) * (2 if accumulate else 1), "hbm"), _ptr(dg),
                   _ptr(a), _ptr(v), _ptr(da), _ptr(dv), int(accumulate), b, n, d)

    def softmax_agg_bwd_bwd(self, ua, uv, dg, a, v, g_dg, g_a, g_v):
        b, n, d = v.shape
        self._call("dg_softmax_agg_bwd_bwd", ("softmax_agg_bwd_bwd", 0, _nbytes(ua, a, g_a), "hbm"), _ptr(ua), _ptr(uv),
                   _ptr(dg), _ptr(a), _ptr(v), _ptr(g_dg), _ptr(g_a), _ptr(g_v), b, n, d)


def _mlp_fwd(self, x, w1, b1, w2, b2, gamma, beta, out, eps, workspace):
    r, d = x.shape
    h = w1.shape[0]
    meta = (f"mlp_fwd[R={r},H={h},fused]", 4 * r * d * h, _nbytes(x, out), "hbm", _nbytes(x, out))
    self._call("dg_mlp_fwd", meta, _ptr(x), _ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2), _ptr(gamma), _ptr(beta), _ptr(out),
               r, d, h, eps, _ptr(workspace), workspace.numel())


def _attn_scores_fwd(self, q, k, v, e, c, a, g, stats=None):
    b, n, d = q.shape
    sm, si = stats if stats is not None else (None, None)
    self._call("dg_attn_scores_fwd", ("attn_scores_fwd[fused]" if a is not None else "attn_scores_fwd[stats only]", 0, _nbytes(e, a), "hbm"), _ptr(q), _ptr(k), _ptr(v), _ptr(e), c,
               _ptr(a), _ptr(g), _ptr(sm), _ptr(si), b, n, d)


def _attn_scores_bwd(self, dg, da_in, q, k, v, e, c, de, dq, dk, dv, stats=None, scores_bf16=False, accumulate_de=False):
    b, n, d = q.shape
    sm, si, g = stats if stats is not None else (None, None, None)
    de16 = de.dtype == torch.bfloat16
    da16 = da_in is not None and da_in.dtype == torch.bfloat16
    self._call("dg_attn_scores_bwd", ("attn_scores_bwd[fused%s%s]" % (",de16" if de16 else "", ",da16" if da16 else ""), 0, _nbytes(e, da_in, de), "hbm"), _ptr(dg),
               _ptr(da_in), _ptr(q), _ptr(k), _ptr(v), _ptr(e), c, _ptr(sm), _ptr(si), _ptr(g), _ptr(de), _ptr(dq), _ptr(dk), _ptr(dv),
               b, n, d, int(de16) | (2 if scores_bf16 else 0) | (4 if da16 else 0) | (8 if accumulate_de else 0))


def _softmax_agg16_fwd(self, a16, v, g, stats=None):
    b, n, d = v.shape
    sm, si = stats if stats is not None else (None, None)
    self._call("dg_softmax_agg16_fwd", ("softmax_agg16_fwd", 0, _nbytes(a16), "hbm"), _ptr(a16), _ptr(v), _ptr(g), _ptr(sm), _ptr(si),
               b, n, d)


def _attn_edge_fwd(self, y, q, k, we, be, woe, boe, gamma, beta, c, out, a16, e_out, z_out, eps, workspace):
    b, n, d = q.shape
    r = b * n * n
    tag = ("+a16" if a16 is not None else "") + ("+e" if e_out is not None else "") + ("+z" if z_out is not None else "")
    meta = (f"attn_edge_fwd[fused{tag}]", 4 * r * d * d, _nbytes(y, out, a16, e_out, z_out), "hbm", _nbytes(y, out))
    self._call("dg_attn_edge_fwd", meta, _ptr(y), _ptr(q), _ptr(k), _ptr(we), _ptr(be), _ptr(woe), _ptr(boe), _ptr(gamma), _ptr(beta),
               c, _ptr(out), _ptr(a16), _ptr(e_out), _ptr(z_out), b, n, d, eps, _ptr(workspace), workspace.numel())


def _mlp_bwd_ln(self, x, dout, w1, b1, w2, b2, gamma, dz, h16, dgamma, dbeta, eps, workspace, mask=None):
    r, d = x.shape
    h = w1.shape[0]
    tag = ("" if h16 is not None else ",no h") + (",mask" if mask is not None else "")
    meta = (f"mlp_bwd_ln[R={r},H={h},fused{tag}]", 4 * r * d * h, _nbytes(x, dout, dz, h16, mask), "hbm", _nbytes(x, dout, dz))
    self._call("dg_mlp_bwd_ln", meta, _ptr(x), _ptr(dout), _ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2), _ptr(gamma), _ptr(dz),
               _ptr(h16), _ptr(mask), _ptr(dgamma), _ptr(dbeta), r, d, h, eps, _ptr(workspace), workspace.numel())


def _mlp_bwd_dgrad(self, dz, h16, w1, w2, dx, dh16, workspace, mask=None):
    r, d = dz.shape
    h = w1.shape[0]
    tag = (",mask" if mask is not None else "") + ("" if dh16 is not None else ",no dh")
    gate = mask if mask is not None else h16
    meta = (f"mlp_bwd_dgrad[R={r},H={h},fused{tag}]", 4 * r * d * h, _nbytes(dz, gate, dx, dh16), "hbm", _nbytes(dz, dx))
    self._call("dg_mlp_bwd_dgrad", meta, _ptr(dz), _ptr(None if mask is not None else h16), _ptr(mask), _ptr(w1), _ptr(w2), _ptr(dx),
               _ptr(dh16), r, d, h, _ptr(workspace), workspace.numel())


def _symmetrize(self, e, out):
    b, n, _, d = e.shape
    self._call("dg_symmetrize", ("symmetrize", 0, _nbytes(e, e, out), "hbm"), _ptr(e), _ptr(out), b, n, d)


def _label2onehot(self, labels, out, classes):
    self._call("dg_label2onehot", ("label2onehot", 0, _nbytes(labels, out), "hbm"), _ptr(labels), labels.element_size(), _ptr(out),
               labels.numel(), classes)


def _argmax_last(self, x, out):
    c = x.shape[-1]
    self._call("dg_argmax_last", ("argmax_last", 0, _nbytes(x, out), "hbm"), _ptr(x), _ptr(out), x.numel() // c, c)


def _embed_labels_fwd(self, labels, lut, y, n, sym):
    self._call("dg_embed_labels_fwd", ("embed_labels_fwd", 0, _nbytes(labels, y), "hbm"), _ptr(labels), labels.element_size(), _ptr(lut),
               _ptr(y), labels.numel(), n, lut.shape[0], lut.shape[1], int(sym))


def _embed_labels_bwd(self, labels, dy, dlut, n, sym):
    self._call("dg_embed_labels_bwd", ("embed_labels_bwd", 0, _nbytes(labels, dy), "hbm"), _ptr(labels), labels.element_size(), _ptr(dy),
               _ptr(dlut), labels.numel(), n, dlut.shape[0], dlut.shape[1], int(sym))


def _gp_interp(self, labels, fake, eps, out, rows_per_mol):
    self._call("dg_gp_interp", ("gp_interp", 0, _nbytes(labels, fake, out), "hbm"), _ptr(labels), labels.element_size(), _ptr(fake),
               _ptr(eps), _ptr(out), labels.numel(), rows_per_mol, fake.shape[-1])


def _gp_penalty(self, g_node, g_edge, penalty, coef, scratch):
    b = g_node.shape[0]
    self._call("dg_gp_penalty", ("gp_penalty", 0, _nbytes(g_node, g_edge), "hbm"), _ptr(g_node), _ptr(g_edge), _ptr(penalty), _ptr(coef),
               _ptr(scratch), b, g_node.numel() // b, g_edge.numel() // b)


def _gp_penalty_bwd(self, g, coef, upstream, out):
    b = g.shape[0]
    self._call("dg_gp_penalty_bwd", ("gp_penalty_bwd", 0, _nbytes(g, out), "hbm"), _ptr(g), _ptr(coef), _ptr(upstream), _ptr(out), b,
               g.numel() // b)


def _readout_argmax(self, x, w, bias, logits, idx):
    rows = x.numel() // x.shape[-1]
    self._call("dg_readout_argmax", ("readout_argmax", 2 * rows * x.shape[-1] * w.shape[0], _nbytes(x, logits, idx), "hbm"), _ptr(x), _ptr(w),
               _ptr(bias), _ptr(logits), _ptr(idx), idx.element_size() if idx is not None else 8, rows, x.shape[-1], w.shape[0])


def _adamw_flat(self, p, g, m, v, segs, nseg, lr, beta1, beta2, eps, wd):
    self._call("dg_adamw_flat", ("adamw_flat", 0, _nbytes(p, g, m, v) + _nbytes(p, m, v), "hbm"), _ptr(p), _ptr(g), _ptr(m), _ptr(v),
               _ptr(segs), nseg, lr, beta1, beta2, eps, wd)


def _to_dense_adj(self, edge_index, batch, edge_attr, adj, cum):
    b, n = adj.shape[0], adj.shape[1]
    self._call("dg_to_dense_adj", ("to_dense_adj", 0, _nbytes(edge_index, batch, edge_attr, adj), "hbm"), _ptr(edge_index), _ptr(batch),
               _ptr(edge_attr), _ptr(adj), _ptr(cum), edge_index.shape[1], batch.numel(), b, n)


def _narrow_labels(self, adj, out, classes):
    self._call("dg_narrow_labels", ("narrow_labels", 0, _nbytes(adj, out), "hbm"), _ptr(adj), _ptr(out), adj.numel(), classes)


def _pack_bits(self, vecs, bits, cnt):
    self._call("dg_pack_bits", ("pack_bits", 0, _nbytes(vecs, bits), "hbm"), _ptr(vecs), vecs.element_size(), _ptr(bits), _ptr(cnt),
               vecs.shape[0], vecs.shape[1])


def _tanimoto_agg(self, sbits, scnt, gbits, gcnt, agg, p, out_max, out_sum):
    self._call("dg_tanimoto_agg", ("tanimoto_agg", 0, _nbytes(sbits, gbits), "hbm"), _ptr(sbits), _ptr(scnt), sbits.shape[0], _ptr(gbits),
               _ptr(gcnt), gbits.shape[0], sbits.shape[1], agg, p, _ptr(out_max), _ptr(out_sum))


def _label_error(self, clear=True):
    return self.lib.dg_label_error(int(clear))


CudaBackend.symmetrize = _symmetrize
CudaBackend.embed_labels_fwd = _embed_labels_fwd
CudaBackend.embed_labels_bwd = _embed_labels_bwd
CudaBackend.gp_interp = _gp_interp
CudaBackend.gp_penalty = _gp_penalty
CudaBackend.gp_penalty_bwd = _gp_penalty_bwd
CudaBackend.readout_argmax = _readout_argmax
CudaBackend.adamw_flat = _adamw_flat
CudaBackend.label_error = _label_error
CudaBackend.to_dense_adj = _to_dense_adj
CudaBackend.narrow_labels = _narrow_labels
CudaBackend.pack_bits = _pack_bits
CudaBackend.tanimoto_agg = _tanimoto_agg
CudaBackend.label2onehot = _label2onehot
CudaBackend.argmax_last = _argmax_last
CudaBackend.mlp_fwd = _mlp_fwd
CudaBackend.mlp_bwd_ln = _mlp_bwd_ln
CudaBackend.mlp_bwd_dgrad = _mlp_bwd_dgrad
CudaBackend.attn_scores_fwd = _attn_scores_fwd
CudaBackend.attn_scores_bwd = _attn_scores_bwd
CudaBackend.softmax_agg16_fwd = _softmax_agg16_fwd
CudaBackend.attn_edge_fwd = _attn_edge_fwd


def cuda_backend() -> CudaBackend:
    global _backend
    if _backend is None:
        if not torch.cuda.is_available():
            raise RuntimeError("druggen_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        _backend = CudaBackend(load())
    return _backend
