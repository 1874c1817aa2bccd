"""A bounds check of the block-level entry points and of their Python callers, WITHOUT a GPU.

block.py's ``_native_forward`` / ``_native_backward`` / ``_native_backward_backward`` allocate every buffer of a block call and
hand the library a table of pointers; csrc/block.cu then addresses scratch arenas by slot arithmetic and reuses slots for
tensors of different widths (h / dh [rows, H] bf16, then dE [rows, D] bf16 in the same buffer, ...).  A buffer sized for the
wrong case is silent memory corruption on the device.  Here the wrappers run on CPU tensors with the library in its dry-run
trace (nothing is launched, every kernel entry point records its arguments), and every pointer a launch would touch is checked
together with its extent -- derived from that launch's own shape arguments -- against the tensors the wrapper allocated.
"""
import os

import pytest
import torch

from druggen_b200 import _lib, block
from druggen_b200 import kernels as K

pytestmark = pytest.mark.skipif(not os.path.exists(_lib.LIB_PATH), reason="library not built (run __graft_entry__.build())")
D, HEADS = 128, 8


def extents(name, a):
    """[(pointer, bytes)] a launch touches; ``a``: the recorded arguments (ints for p: / i:, floats for f:)."""
    if name in ("dg_add_ln_fwd",):
        r, d = a[5], a[6]
        return [(a[0], r * d * 4), (a[1], r * d * 4), (a[2], d * 4), (a[3], d * 4), (a[4], r * d * 4)]
    if name == "dg_add_ln_bwd":
        r, d = a[7], a[8]
        return [(a[i], r * d * 4) for i in (0, 1, 2, 4)] + [(a[i], d * 4) for i in (3, 5, 6)]
    if name == "dg_add_ln_bwd_bwd":
        r, d = a[10], a[11]
        return [(a[i], r * d * 4) for i in (0, 3, 4, 5, 7, 8)] + [(a[i], d * 4) for i in (1, 2, 6, 9)]
    if name == "dg_rows_gemm":
        r, k, n, flags = a[8], a[9], a[10], a[12]
        return [(a[0], r * k * (2 if flags & 1 else 4)), (a[1], k * n * 4), (a[3], n * 4), (a[5], r * n * (2 if flags & 4 else 4)),
                (a[6], r * n * 4), (a[7], r * n * (2 if flags & 2 else 4))]
    if name == "dg_gemm_tn":
        r, m, n, flags = a[4], a[5], a[6], a[8]
        return [(a[0], r * m * (2 if flags & 1 else 4)), (a[1], r * n * (2 if flags & 2 else 4)), (a[2], m * n * 4), (a[3], m * 4)]
    if name == "dg_attn_edge_fwd":
        b, n, d = a[14], a[15], a[16]
        e, nd = b * n * n * d, b * n * d
        return ([(a[i], e * 4) for i in (0, 10, 12, 13)] + [(a[11], e * 2), (a[1], nd * 4), (a[2], nd * 4), (a[3], d * d * 4), (a[5], d * d * 4)] +
                [(a[i], d * 4) for i in (4, 6, 7, 8)] + [(a[18], 65536)])
    if name == "dg_softmax_agg16_fwd":
        b, n, d = a[5], a[6], a[7]
        return [(a[0], b * n * n * d * 2)] + [(a[i], b * n * d * 4) for i in (1, 2, 3, 4)]
    if name == "dg_attn_scores_fwd":
        b, n, d = a[9], a[10], a[11]
        return [(a[i], b * n * d * 4) for i in (0, 1, 2, 6, 7, 8)] + [(a[i], b * n * n * d * 4) for i in (3, 5)]
    if name == "dg_attn_scores_bwd":
        b, n, d, flags = a[14], a[15], a[16], a[17]
        e = b * n * n * d
        return ([(a[i], b * n * d * 4) for i in (0, 2, 3, 4, 7, 8, 9, 11, 12, 13)] +
                [(a[1], e * (2 if flags & 4 else 4)), (a[5], e * 4), (a[10], e * (2 if flags & 1 else 4))])
    if name == "dg_mlp_fwd":
        r, d, h = a[8], a[9], a[10]
        return [(a[0], r * d * 4), (a[7], r * d * 4), (a[1], h * d * 4), (a[3], h * d * 4), (a[2], h * 4), (a[4], d * 4), (a[5], d * 4), (a[6], d * 4),
                (a[12], 2 * (h // 128) * 32768)]
    if name == "dg_mlp_bwd_ln":
        r, d, h = a[12], a[13], a[14]
        return ([(a[i], r * d * 4) for i in (0, 1, 7)] + [(a[2], h * d * 4), (a[4], h * d * 4), (a[3], h * 4), (a[5], d * 4), (a[6], d * 4),
                (a[8], r * h * 2), (a[9], r * (h // 64) * 8), (a[10], d * 4), (a[11], d * 4), (a[16], 2 * (h // 128) * 32768)])
    if name == "dg_mlp_bwd_dgrad":
        r, d, h = a[7], a[8], a[9]
        return [(a[0], r * d * 4), (a[5], r * d * 4), (a[1], r * h * 2), (a[6], r * h * 2), (a[2], r * (h // 64) * 8), (a[3], h * d * 4), (a[4], h * d * 4),
                (a[10], 2 * (h // 128) * 32768)]
    if name == "dg_modulate_bwd":
        b, n, d = a[8], a[9], a[10]
        return [(a[i], b * n * n * d * 4) for i in (0, 3, 7)] + [(a[i], b * n * d * 4) for i in (1, 2, 5, 6)]
    if name == "dg_modulate_bwd_bwd":
        b, n, d = a[12], a[13], a[14]
        return [(a[i], b * n * n * d * 4) for i in (2, 3, 6, 8, 11)] + [(a[i], b * n * d * 4) for i in (0, 1, 4, 5, 9, 10)]
    if name == "dg_softmax_agg_bwd":
        b, n, d = a[6], a[7], a[8]
        return [(a[i], b * n * n * d * 4) for i in (1, 3)] + [(a[i], b * n * d * 4) for i in (0, 2, 4)]
    if name == "dg_softmax_agg_bwd_bwd":
        b, n, d = a[8], a[9], a[10]
        return [(a[i], b * n * n * d * 4) for i in (0, 3, 6)] + [(a[i], b * n * d * 4) for i in (1, 2, 4, 5, 7)]
    if name == "memset0":
        return [(a[0], a[1])]
    if name == "transpose":
        return [(a[0], a[2] * a[3] * 4), (a[1], a[2] * a[3] * 4)]
    if name == "add3":
        return [(a[i], a[6] * 4) for i in range(6)]
    raise AssertionError("no extent table for " + name)


class TraceBackend(_lib.CudaBackend):
    """The real launch table with the library in its dry-run trace: CPU tensors, no stream, nothing launched."""

    def __init__(self):
        super().__init__(_lib.load())
        self.tensors = []                                  # every tensor handed to a block-level call

    def native_blocks(self):
        return True

    def _native(self, name, *args):
        rc = getattr(self.lib, name)(*args, None)
        if rc != 0:
            raise RuntimeError(f"{name} rejected: {self.lib.dg_last_error().decode()}")

    def _remember(self, io, params, grads, ws):
        self.tensors += [t for t in list(io.values()) + list(params) + list(grads or []) + [ws] if t is not None]

    def block_fwd(self, io, params, b, n, d, h, heads, flags, eps, ws):
        self._remember(io, params, None, ws)
        super().block_fwd(io, params, b, n, d, h, heads, flags, eps, ws)

    def block_bwd(self, io, params, grads, b, n, d, h, heads, flags, eps, ws):
        self._remember(io, params, grads, ws)
        super().block_bwd(io, params, grads, b, n, d, h, heads, flags, eps, ws)

    def block_bwd_bwd(self, io, params, grads, b, n, d, h, heads, flags, eps, ws):
        self._remember(io, params, grads, ws)
        super().block_bwd_bwd(io, params, grads, b, n, d, h, heads, flags, eps, ws)

    def encoder_fwd(self, x, y, x_out, y_out, params, depth, scratch, b, n, d, h, heads, last_edge_out, eps, ws):
        self._remember(dict(scratch, _x=x, _y=y, _xo=x_out, _yo=y_out), params, None, ws)
        super().encoder_fwd(x, y, x_out, y_out, params, depth, scratch, b, n, d, h, heads, last_edge_out, eps, ws)


@pytest.fixture()
def traced(monkeypatch):
    be = TraceBackend()
    monkeypatch.setattr(_lib, "_backend", be)
    monkeypatch.setattr(_lib, "cuda_backend", lambda: be)
    monkeypatch.setattr(K, "_chk", lambda *a, **k: None)
    monkeypatch.setattr(K, "_chk_buffers", lambda *a, **k: None)
    monkeypatch.setattr(K, "native_block_available", lambda *a, **k: True)
    monkeypatch.setattr(K, "_precision", "bf16")

    def aligned_ws(w1):            # (the CUDA allocator hands out 512-byte-aligned blocks; the CPU one 64-byte-aligned ones)
        nbytes = 2 * (w1.shape[0] // 128) * 32768
        raw = torch.empty(nbytes + 128, dtype=torch.uint8)
        off = (-raw.data_ptr()) % 128
        return raw[off:off + nbytes]
    monkeypatch.setattr(K, "_mlp_ws", aligned_ws)
    be.lib.dg_debug_trace(1)
    yield be
    be.lib.dg_debug_trace(0)


def read_trace(be):
    import ctypes as C
    buf = C.create_string_buffer(1 << 20)
    be.lib.dg_debug_trace_read(buf, len(buf))
    out = []
    for line in buf.value.decode().strip().splitlines():
        name, *args = line.split(" ")
        vals = []
        for tok in args:
            kind, v = tok.split(":", 1)
            vals.append(int(v, 16) if kind == "p" else (int(v) if kind == "i" else float(v)))
        out.append((name, vals))
    return out


def check_bounds(be):
    prog = read_trace(be)
    assert prog, "nothing was traced"
    ranges = sorted({(t.data_ptr(), t.data_ptr() + t.numel() * t.element_size()) for t in be.tensors if t.numel()})
    checked = 0
    for name, vals in prog:
        for ptr, nbytes in extents(name, vals):
            if ptr == 0:
                continue
            assert nbytes > 0, (name, vals)
            inside = any(lo <= ptr and ptr + nbytes <= hi for lo, hi in ranges)
            assert inside, "%s touches [%x, +%d) which is not inside any buffer of the call: %s" % (name, ptr, nbytes, vals)
            checked += 1
    be.tensors = []
    return len(prog), checked


def make_params(hid):
    g = torch.Generator().manual_seed(0)
    shapes = {"fc1.weight": (hid, D), "fc1.bias": (hid,), "fc2.weight": (D, hid)}
    params = []
    for nm in block.BLOCK_PARAM_NAMES:
        shape = next((s for k, s in shapes.items() if nm.endswith(k)), (D, D) if nm.endswith("weight") and ".ln" not in nm and not nm.startswith("ln") else (D,))
        params.append(torch.randn(*shape, generator=g))
    return params


def data(b, n):
    return torch.randn(b, n, D), torch.randn(b, n, n, D)


@pytest.mark.parametrize("hid", [384, 128])
@pytest.mark.parametrize("b,n", [(3, 9), (2, 45), (1, 4)])
def test_forward_buffers(traced, b, n, hid):
    params = make_params(hid)
    x, y = data(b, n)
    for edge_out, want_stats, want_saved in ((True, True, True), (True, True, False), (True, False, None), (False, True, False), (False, False, None)):
        block.block_forward_nograd(x, y, params, HEADS, edge_out, want_stats=want_stats, want_saved=want_saved)
        launches, checked = check_bounds(traced)
        assert launches >= 9 and checked >= 4 * launches


@pytest.mark.parametrize("hid", [384, 256])
@pytest.mark.parametrize("b,n", [(3, 9), (2, 45)])
def test_backward_buffers(traced, b, n, hid):
    params = make_params(hid)
    x, y = data(b, n)
    dxo, dyo = data(b, n)
    for edge_out, want_params, kept, have_dxo in ((True, True, True, True), (True, True, False, True), (True, False, True, True),
                                                  (True, False, False, True), (False, True, False, True), (False, False, False, True),
                                                  (True, True, True, False), (True, False, False, False)):
        xo, yo, stats, saved = block.block_forward_nograd(x, y, params, HEADS, edge_out, want_stats=True, want_saved=kept and edge_out)
        check_bounds(traced)
        block.block_backward(x, y, dxo if have_dxo else None, dyo if edge_out else None, params, HEADS, edge_out, want_params, stats,
                             saved if kept else None)
        launches, checked = check_bounds(traced)
        assert launches >= 15 and checked >= 4 * launches, (launches, checked)


@pytest.mark.parametrize("hid", [384, 128])
@pytest.mark.parametrize("b,n", [(3, 9), (2, 45)])
def test_second_order_buffers(traced, b, n, hid):
    params = make_params(hid)
    x, y = data(b, n)
    dxo, dyo = data(b, n)
    ux, uy = data(b, n)
    for edge_out, kept, have_uy in ((True, True, True), (True, False, True), (False, False, True), (True, True, False)):
        saved = None
        if kept:
            saved = block.block_forward_nograd(x, y, params, HEADS, edge_out, want_stats=True, want_saved=True)[3]
            check_bounds(traced)
        block.block_backward_backward(x, y, dxo, dyo if edge_out else None, ux, uy if have_uy else None, params, HEADS, edge_out, saved)
        launches, checked = check_bounds(traced)
        assert launches >= 50 and checked >= 4 * launches, (launches, checked)


def test_encoder_forward_buffers(traced, monkeypatch):
    monkeypatch.setitem(block._GRAPH, "on", False)                     # (the CUDA-graph replay needs a device)
    x, y = data(3, 9)
    for depth in (1, 2, 3, 4):
        for last_edge_out in (True, False):
            blocks = [make_params(384) for _ in range(depth)]
            block.encoder_forward_nograd(x, y, blocks, HEADS, last_edge_out)
            launches, checked = check_bounds(traced)
            assert launches == 10 * depth - (0 if last_edge_out else 1)


# ---- the C launch programs against block.py's launch lists, structurally -----------------------------------------------
# Both run on the same CPU inputs through the dry-run trace.  Pointers are replaced by value numbers: the call's inputs and
# parameters keep their names, every other buffer gets the number of the launch output that defined it (a slot that the C
# side reuses gets a new number with every write, like a fresh tensor on the Python side), buffers first seen as an
# accumulation target count as fresh zeros.  The two programs must then be IDENTICAL: same launches, same order, same scalars,
# same data flow.
WS_POS = {"dg_attn_edge_fwd": 18, "dg_mlp_fwd": 12, "dg_mlp_bwd_ln": 16, "dg_mlp_bwd_dgrad": 10}
PTR_POS = {   # kernel -> (read positions, written positions, accumulated positions)
    "dg_add_ln_fwd": ((0, 1, 2, 3), (4,), ()),
    "dg_add_ln_bwd": ((0, 1, 2, 3), (4,), (5, 6)),
    "dg_add_ln_bwd_bwd": ((0, 1, 2, 3, 4, 5, 6), (7, 8), (9,)),
    "dg_rows_gemm": ((0, 1, 3, 5, 6), (7,), ()),
    "dg_gemm_tn": ((0, 1), (), (2, 3)),
    "dg_attn_edge_fwd": ((0, 1, 2, 3, 4, 5, 6, 7, 8), (10, 11, 12, 13), ()),
    "dg_softmax_agg16_fwd": ((0, 1), (2, 3, 4), ()),
    "dg_attn_scores_fwd": ((0, 1, 2, 3), (5, 6, 7, 8), ()),
    "dg_attn_scores_bwd": ((0, 1, 2, 3, 4, 5, 7, 8, 9), (10,), (11, 12, 13)),
    "dg_mlp_fwd": ((0, 1, 2, 3, 4, 5, 6), (7,), ()),
    "dg_mlp_bwd_ln": ((0, 1, 2, 3, 4, 5, 6), (7, 8, 9), (10, 11)),
    "dg_mlp_bwd_dgrad": ((0, 1, 2, 3, 4), (5, 6), ()),
    "dg_modulate_bwd": ((0, 1, 2, 3), (5, 7), (6,)),
    "dg_modulate_bwd_bwd": ((0, 1, 2, 3, 4, 5, 6), (8, 9, 11), (10,)),
    "dg_softmax_agg_bwd": ((0, 1, 2), (3,), (4,)),
    "dg_softmax_agg_bwd_bwd": ((0, 1, 2, 3, 4), (5, 6), (7,)),
}
COND = {"dg_attn_scores_bwd": (10, 17, 8), "dg_add_ln_bwd": (4, 10, 1), "dg_softmax_agg_bwd": (3, 5, 1)}


def canonical(prog, named, scratch=()):
    """Value-numbered form of a traced program.  ``named``: address -> name of the call's inputs / parameters; ``scratch``:
    address ranges whose contents nobody reads (the dgrad-only passes' throw-away LayerNorm affine gradients)."""
    val, out, fresh = dict(named), [], [0]

    def new(prefix):
        fresh[0] += 1
        return "%s%d" % (prefix, fresh[0])

    for name, a in prog:
        if name in ("memset0", "transpose"):            # the C side's own fills / weight transposes: the Python side gets fresh
            lo, hi = (a[0], a[0] + a[1]) if name == "memset0" else (a[1], a[1] + a[2] * a[3] * 4)     # zeros / copies from torch
            for addr in [k for k in val if lo <= k < hi and k not in named]:
                del val[addr]
            continue
        if name == "add3":                              # c[q] += dq2 ...: in place on both sides (torch's add_ is not traced)
            continue
        reads, writes, accs = PTR_POS[name]
        accs = list(accs)
        if name in COND and a[COND[name][1]] & COND[name][2]:
            accs.append(COND[name][0])
        ptrs = set(reads) | set(writes) | set(accs) | {WS_POS.get(name, -1)}
        row = [name]
        for i, v in enumerate(a):
            if i == WS_POS.get(name):
                continue
            if i not in ptrs:
                row.append(v)
            elif v == 0:
                row.append("0")
            elif any(lo <= v < hi for lo, hi in scratch):
                row.append("scratch")
            elif i in reads or i in accs:
                if v not in val:
                    val[v] = new("in")                  # first seen as an input: a fresh (zeroed / copied) tensor
                row.append(val[v])
                if i in accs:
                    val[v] = new("v")
            else:
                row.append("->")
        for i in writes:                                # outputs are numbered after the reads: a launch may overwrite its input slot
            if i < len(a) and a[i] != 0 and i not in accs and not any(lo <= a[i] < hi for lo, hi in scratch):
                assert a[i] not in named, (name, "writes an input of the call")
                val[a[i]] = new("v")
        out.append(tuple(row))
    return out


class PyTraceBackend(TraceBackend):
    """block.py's launch-by-launch lists through the same dry run (no stream, nothing launched)."""

    def _launch(self, name, key, flops, nbytes, alg, bound, timed, args):
        rc = getattr(self.lib, name)(*args, None)
        if rc != 0:
            raise RuntimeError(f"{name} rejected: {self.lib.dg_last_error().decode()}")


@pytest.fixture()
def both(monkeypatch):
    """-> run(fn): the traced programs of fn() with the block-level entry points and with block.py's own lists."""
    be = PyTraceBackend()
    monkeypatch.setattr(_lib, "_backend", be)
    monkeypatch.setattr(_lib, "cuda_backend", lambda: be)
    monkeypatch.setattr(K, "_chk", lambda *a, **k: None)
    monkeypatch.setattr(K, "_chk_buffers", lambda *a, **k: None)
    monkeypatch.setattr(K, "_precision", "bf16")
    monkeypatch.setattr(K, "attn_chain_available", lambda *a, **k: True)
    ws_keep = []

    def aligned_ws(w1):
        nbytes = 2 * (w1.shape[0] // 128) * 32768
        raw = torch.empty(nbytes + 128, dtype=torch.uint8)
        ws_keep.append(raw)
        off = (-raw.data_ptr()) % 128
        return raw[off:off + nbytes]
    monkeypatch.setattr(K, "_mlp_ws", aligned_ws)
    # every tensor of a run stays alive until both programs are compared: an address then names ONE logical tensor
    keep = []
    for fn_name in ("empty", "zeros", "empty_like", "zeros_like"):
        orig = getattr(torch, fn_name)

        def alloc(*a, _orig=orig, **k):
            t = _orig(*a, **k)
            keep.append(t)
            return t
        monkeypatch.setattr(torch, fn_name, alloc)

    orig_contig = torch.Tensor.contiguous

    def contig(self, *a, **k):                     # (w.t().contiguous(): the transposed weights of the second-order pass)
        t = orig_contig(self, *a, **k)
        keep.append(t)
        return t
    monkeypatch.setattr(torch.Tensor, "contiguous", contig)

    def run(fn, named_tensors):
        named = {}
        for nm, t in named_tensors.items():
            if t is not None:
                named[t.data_ptr()] = nm
        progs = []
        for native in (True, False):
            monkeypatch.setattr(K, "native_block_available", lambda *a, _n=native, **k: _n)
            be.lib.dg_debug_trace(1)
            try:
                res = fn()
                keep.append(res)
                prog = read_trace(be)
            finally:
                be.lib.dg_debug_trace(0)
            scratch = [(t.data_ptr(), t.data_ptr() + 1024) for t in be.tensors if t.numel() == 2 * D and t.dtype == torch.float32] if native else []
            be.tensors = []
            progs.append(canonical(prog, named, scratch))
        return progs
    return run


def _scrub_scratch(prog):
    """The affine gradients of dg_add_ln_bwd are the one place where the two sides are wired differently on purpose: block.py hands
    every launch fresh zeroed dgamma / dbeta tensors and adds them into the parameter cotangents with torch (or drops them, on
    dgrad-only passes); the library accumulates straight into the cotangent table (or into one scratch vector).  Compared as a
    placeholder here; where they end up is pinned by tests/golden/native_block_programs.json and, numerically, on the GPU."""
    out = []
    for row in prog:
        if row[0] == "dg_add_ln_bwd":
            row = list(row)
            row[6] = row[7] = "affine"                     # (positions shift by one: row[0] is the kernel name)
            row = tuple(row)
        out.append(row)
    return out


def _renumber(prog):
    """Value numbers in order of first appearance (dropping the throw-away tensors above shifts the counters)."""
    m, out = {}, []
    for row in prog:
        new = []
        for v in row:
            if isinstance(v, str) and (v[:2] == "in" or v[:1] == "v") and v[-1].isdigit() and v not in ("scratch",):
                v = m.setdefault(v, "%s#%d" % ("in" if v.startswith("in") else "v", len(m)))
            new.append(v)
        out.append(tuple(new))
    return out


def assert_same(progs):
    a, b = (_renumber(_scrub_scratch(p)) for p in progs)
    assert len(a) == len(b), (len(a), len(b), [r[0] for r in a], [r[0] for r in b])
    for i, (ra, rb) in enumerate(zip(a, b)):
        assert ra == rb, "launch %d differs:\n  library : %s\n  block.py: %s" % (i, ra, rb)


@pytest.mark.parametrize("edge_out,want_stats,want_saved", [(True, True, True), (True, True, False), (True, False, None), (False, False, None)])
def test_forward_program_equals_python_list(both, edge_out, want_stats, want_saved):
    params = make_params(384)
    x, y = data(3, 9)
    named = {"X": x, "Y": y, **{"P%d" % i: p for i, p in enumerate(params)}}
    assert_same(both(lambda: block.block_forward_nograd(x, y, params, HEADS, edge_out, want_stats=want_stats, want_saved=want_saved), named))


@pytest.mark.parametrize("edge_out,want_params,have_dxo", [(True, True, True), (True, False, True), (False, True, True), (False, False, True),
                                                           (True, True, False)])
def test_backward_program_equals_python_list(both, edge_out, want_params, have_dxo):
    """(recomputing backward with the forward's statistics: kept intermediates would come from two different forward runs)"""
    params = make_params(384)
    x, y = data(3, 9)
    dxo, dyo = data(3, 9)
    dxo = dxo if have_dxo else None
    dyo = dyo if edge_out else None
    stats = tuple(torch.randn(3, 9, D) for _ in range(3)) if edge_out else None
    named = {"X": x, "Y": y, "DXO": dxo, "DYO": dyo, **{"P%d" % i: p for i, p in enumerate(params)}}
    if stats:
        named.update({"STAT_M": stats[0], "STAT_INV": stats[1], "G": stats[2]})
    assert_same(both(lambda: block.block_backward(x, y, dxo, dyo, params, HEADS, edge_out, want_params, stats, None), named))


@pytest.mark.parametrize("edge_out,kept", [(True, False), (False, False), (True, True)])
def test_second_order_program_equals_python_list(both, edge_out, kept):
    params = make_params(384)
    x, y = data(3, 9)
    dxo, dyo = data(3, 9)
    ux, uy = data(3, 9)
    dyo = dyo if edge_out else None
    named = {"X": x, "Y": y, "DXO": dxo, "DYO": dyo, "UX": ux, "UY": uy, **{"P%d" % i: p for i, p in enumerate(params)}}
    saved = None
    if kept:
        saved = {"x1": torch.randn(27, D), "q": torch.randn(3, 9, D), "k": torch.randn(3, 9, D), "v": torch.randn(3, 9, D),
                 "y3": torch.randn(243, D), "e": torch.randn(243, D), "z4": torch.randn(243, D)}
        named.update({nm.upper(): t for nm, t in saved.items()})
    assert_same(both(lambda: block.block_backward_backward(x, y, dxo, dyo, ux, uy, params, HEADS, edge_out, saved), named))


# ---- a whole GAN step through the dry run -------------------------------------------------------------------------------
def extents_glue(name, a):
    if name == "dg_symmetrize":
        return [(a[0], a[2] * a[3] * a[3] * a[4] * 4), (a[1], a[2] * a[3] * a[3] * a[4] * 4)]
    if name in ("dg_embed_labels_fwd", "dg_embed_labels_bwd"):
        rows, classes, d = a[4], a[6], a[7]
        return [(a[0], rows * a[1]), (a[2], (classes if name.endswith("fwd") else rows) * d * 4), (a[3], (rows if name.endswith("fwd") else classes) * d * 4)]
    if name == "dg_gp_interp":
        rows, per, classes = a[5], a[6], a[7]
        return [(a[0], rows * a[1]), (a[2], rows * classes * 4), (a[3], rows // per * 4), (a[4], rows * classes * 4)]
    if name == "dg_gp_penalty":
        return [(a[0], a[5] * a[6] * 4), (a[1], a[5] * a[7] * 4), (a[2], 4), (a[3], a[5] * 4), (a[4], a[5] * 4)]
    if name == "dg_gp_penalty_bwd":
        return [(a[0], a[4] * a[5] * 4), (a[1], a[4] * 4), (a[2], 4), (a[3], a[4] * a[5] * 4)]
    if name == "dg_readout_argmax":
        rows, d, classes = a[6], a[7], a[8]
        return [(a[0], rows * d * 4), (a[1], classes * d * 4), (a[2], classes * 4), (a[3], rows * classes * 4), (a[4], rows * a[5])]
    if name == "dg_adamw_flat":
        return [(a[i], 4) for i in range(4)] + [(a[4], a[5] * 32)]          # (the segment table describes the four flat buffers)
    if name == "dg_label2onehot":
        return [(a[0], a[3] * a[1]), (a[2], a[3] * a[4] * 4)]
    if name == "dg_argmax_last":
        return [(a[0], a[2] * a[3] * 4), (a[1], a[2] * 8)]
    if name == "dg_gate_mul":
        return [(a[i], a[3] * 4) for i in range(3)]
    if name == "dg_colsum":
        return [(a[0], a[2] * a[3] * 4), (a[1], a[3] * 4)]
    if name == "dg_modulate_fwd":
        b, n, d = a[5], a[6], a[7]
        return [(a[0], b * n * d * 4), (a[1], b * n * d * 4), (a[2], b * n * n * d * 4), (a[4], b * n * n * d * 4)]
    if name == "dg_softmax_agg_fwd":
        b, n, d = a[3], a[4], a[5]
        return [(a[0], b * n * n * d * 4), (a[1], b * n * d * 4), (a[2], b * n * d * 4)]
    return extents(name, a)


@pytest.mark.parametrize("precision,native", [("bf16", True), ("bf16", False), ("bf16x3", False), ("fp32", False)])
def test_whole_gan_step_bounds(monkeypatch, precision, native):
    """One GANTrainer.step (D step with the gradient penalty's double backward, G step, both AdamW updates) on CPU tensors through
    the dry run: every launch of the step -- issued by the block-level entry points or one by one from block.py -- touches only
    memory inside the tensors it was handed, given the shapes it was told."""
    import druggen_b200 as dg
    from druggen_b200 import gan
    be = PyTraceBackend()
    seen = []
    orig_ptr = _lib._ptr

    def ptr(t):
        if t is not None:
            seen.append(t)
        return orig_ptr(t)
    monkeypatch.setattr(_lib, "_ptr", ptr)
    monkeypatch.setattr(_lib, "_backend", be)
    monkeypatch.setattr(_lib, "cuda_backend", lambda: be)

    def chk(*ts, bf16_ok=False):                          # kernels._chk minus "is on a CUDA device"
        for t in ts:
            assert t is None or (t.is_contiguous() and (t.dtype == torch.float32 or (bf16_ok and t.dtype == torch.bfloat16)))
    monkeypatch.setattr(K, "_chk", chk)
    monkeypatch.setattr(K, "_chk_buffers", lambda ts, dev: None)
    monkeypatch.setattr(K, "_chk_labels", lambda l: l.contiguous())
    monkeypatch.setattr(K, "_precision", precision)
    monkeypatch.setenv("DRUGGEN_B200_NATIVE_BLOCK", "1" if native else "0")

    def aligned_ws(w1):
        nbytes = 2 * (w1.shape[0] // 128) * 32768
        raw = torch.empty(nbytes + 128, dtype=torch.uint8)
        off = (-raw.data_ptr()) % 128
        return raw[off:off + nbytes]
    monkeypatch.setattr(K, "_mlp_ws", aligned_ws)
    torch.manual_seed(0)
    n, bsz = 9, 4
    G = dg.Generator("relu", n, 5, 13, 0.0, dim=D, depth=2, heads=HEADS, mlp_ratio=3)
    Dn = dg.Discriminator("relu", n, 5, 13, 0.0, dim=D, depth=2, heads=HEADS, mlp_ratio=3)
    tr = gan.GANTrainer(G, Dn)
    a, x = gan.synthetic_molecules(bsz, n, 13, 5, seed=3, labels=True)
    be.lib.dg_debug_trace(1)
    try:
        tr.step(a, x, a, x)
        prog = read_trace(be)
    finally:
        be.lib.dg_debug_trace(0)
    names = [p[0] for p in prog]
    assert ("add3" in names) == native and names.count("dg_adamw_flat") == 2 and names.count("dg_gp_penalty") == 1
    assert names.count("dg_add_ln_bwd_bwd") == 8                                           # LN1, LN3, LN5 of two D blocks + LN4, LN6 of the first
    assert (names.count("dg_attn_edge_fwd") >= 6) == (precision == "bf16")                # the fused chains are the throughput mode's
    ranges = sorted({(t.data_ptr(), t.data_ptr() + t.numel() * t.element_size()) for t in seen + be.tensors if t.numel()})
    checked = 0
    for name, vals in prog:
        for p, nbytes in extents_glue(name, vals):
            if p == 0 or nbytes == 0:
                continue
            assert any(lo <= p and p + nbytes <= hi for lo, hi in ranges), \
                "%s touches [%x, +%d) which is not inside any tensor it was handed: %s" % (name, p, nbytes, vals)
            checked += 1
    assert checked > 2000


@pytest.mark.parametrize("depth,last_edge_out", [(1, True), (2, True), (3, True), (4, True), (1, False), (2, False), (3, False), (4, False)])
def test_encoder_program_equals_python_list(both, depth, last_edge_out):
    """dg_encoder_fwd (ping-pong buffers between layers, the caller's outputs used as one of them) against block.py's layer loop."""
    blocks = [make_params(384) for _ in range(depth)]
    x, y = data(3, 9)
    named = {"X": x, "Y": y}
    for l, params in enumerate(blocks):
        named.update({"P%d.%d" % (l, i): p for i, p in enumerate(params)})
    assert_same(both(lambda: block.encoder_forward_nograd(x, y, blocks, HEADS, last_edge_out), named))
