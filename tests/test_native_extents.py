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
            captured = []
            orig = traced.encoder_fwd

            def enc(x_, y_, xo, yo, params, depth_, scratch, *a, _orig=orig, _cap=captured):
                traced.tensors += [t for t in [x_, y_, xo, yo] + list(params) + list(scratch.values()) + [a[-1]] if t is not None]
                return _orig(x_, y_, xo, yo, params, depth_, scratch, *a)
            monkeypatch.setattr(traced, "encoder_fwd", enc)
            block.encoder_forward_nograd(x, y, blocks, HEADS, last_edge_out)
            launches, checked = check_bounds(traced)
            assert launches == 10 * depth - (0 if last_edge_out else 1)
            monkeypatch.setattr(traced, "encoder_fwd", orig)
