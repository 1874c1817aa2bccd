#!/usr/bin/env python
"""Symbolic launch programs of the block-level entry points (dg_block_fwd / dg_block_bwd / dg_block_bwd_bwd), listed WITHOUT a GPU.

With the library's dry-run trace on (dg_debug_trace) every kernel entry point records its name and arguments and returns; the
entry points are called with fake, unique addresses per buffer slot / parameter, and the recorded pointers are mapped back to
their names: `dg_add_ln_fwd(X, 0, P.ln1.weight, P.ln1.bias, X1, 405, 128, 1e-05)`.

    python tools/trace_block.py                 # print every program
    python tools/trace_block.py --write         # regenerate tests/golden/native_block_programs.json
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from druggen_b200 import _lib  # noqa: E402
from druggen_b200.block import BLOCK_PARAM_NAMES  # noqa: E402

B, N, D, H, HEADS = 5, 9, 128, 384, 8
GOLDEN = os.path.join(ROOT, "tests", "golden", "native_block_programs.json")
SLOT_BASE, PARAM_BASE, GRAD_BASE, WS = 0x10_0000_0000, 0x20_0000_0000, 0x30_0000_0000, 0x40_0000_0000
STRIDE = 0x1000_0000                                   # 256 MB of address space per buffer: offsets inside a buffer stay attributable


# node-sized scratch slots of dg_block_bwd_bwd inside N_ARENA, in the order of csrc/block.cu's enum
ARENA = ("t5", "m_n", "dx3", "dz3", "dg", "dv", "dq", "dk", "p0", "p1", "c_dx1", "c_dq", "c_dk", "c_dv", "c_q", "c_k", "c_v", "c_dg", "c_dz3",
         "c_dx3", "c_z3", "c_t5", "c_mn", "c_x3", "c_g", "dq2", "dk2", "dv2")
assert len(ARENA) == _lib.BB_NODE_SLOTS


def names():
    m = {0: "0", WS: "WS"}
    for i, s in enumerate(_lib.BLK_SLOTS):
        m[SLOT_BASE + i * STRIDE] = s
    for i, s in enumerate(BLOCK_PARAM_NAMES):
        m[PARAM_BASE + i * STRIDE] = "P." + s
        m[GRAD_BASE + i * STRIDE] = "G." + s
    return m


def decode(text, bn_elems):
    m = names()
    out = []
    for line in text.strip().splitlines():
        name, *args = line.split(" ")
        dec = []
        for a in args:
            kind, val = a.split(":", 1)
            if kind != "p":
                dec.append(val)
                continue
            addr = int(val, 16)
            base = addr - (addr % STRIDE) if addr >= SLOT_BASE else addr
            off = addr - base
            nm = m.get(base, hex(addr))
            if nm == "N_ARENA" and off % (bn_elems * 4) == 0:
                nm = "N." + ARENA[off // (bn_elems * 4)]           # node-arena slot
            elif off:                                   # a sub-buffer: the second transposed weight, the second scratch vector
                nm += "+%dB" % off
            dec.append(nm)
        out.append("%s(%s)" % (name, ", ".join(dec)))
    return out


def run(entry, slots, flags, grads=True, dead_edge_params=False, second_order=False):
    lib = _lib.load()
    io = (C.c_void_p * len(_lib.BLK_SLOTS))()
    for s in slots:
        io[_lib.BLK[s]] = SLOT_BASE + _lib.BLK[s] * STRIDE
    params = (C.c_void_p * 30)(*[PARAM_BASE + i * STRIDE for i in range(30)])
    gtab = None
    if grads:
        dead = ("attn.out_e.", "ln4.", "mlp2.", "ln6.") if dead_edge_params else ()
        dead += ("ln5.bias", "ln6.bias") if second_order else ()
        gtab = (C.c_void_p * 30)(*[None if (dead and nm.startswith(dead)) else GRAD_BASE + i * STRIDE
                                   for i, nm in enumerate(BLOCK_PARAM_NAMES)])
    lib.dg_debug_trace(1)
    try:
        if entry == "dg_block_fwd":
            rc = lib.dg_block_fwd(io, params, B, N, D, H, HEADS, flags, 1e-5, WS, 2 * 3 * 32768, None)
        else:
            rc = getattr(lib, entry)(io, params, gtab, B, N, D, H, HEADS, flags, 1e-5, WS, 2 * 3 * 32768, None)
        if rc:
            raise RuntimeError(lib.dg_last_error().decode())
        buf = C.create_string_buffer(1 << 20)
        lib.dg_debug_trace_read(buf, len(buf))
    finally:
        lib.dg_debug_trace(0)
    return decode(buf.value.decode(), B * N * D)


FWD_NODE = ["X", "Y", "X_OUT", "X1", "Q", "K", "V", "G", "ON", "X3"]
BWD_COMMON = ["X", "Y", "X1", "Q", "K", "V", "G", "ON", "X3", "STAT_M", "STAT_INV", "E", "DXO", "DX", "DY", "N_DZ", "N_DX3", "N_DZ3", "N_DG",
              "N_DQ", "N_DK", "N_DV", "N_T0", "N_T1", "N_H", "N_MASK", "E_H", "SCRATCH"]
BWD_LIVE = ["DYO", "Y3", "A16", "Z4", "E_A", "E_B", "E_MASK"]
BB_COMMON = ["X", "Y", "DXO", "UX", "UY", "X1", "Q", "K", "V", "G", "STAT_M", "STAT_INV", "ON", "X3", "E", "C_X", "C_Y", "C_DXO", "N_ARENA",
             "N_H", "N_H2", "N_H3", "N_MASK", "ES0", "ES5", "ES6", "ES7", "ES8", "WT", "SCRATCH"]
BB_LIVE = ["DYO", "Y3", "Z4", "C_DYO", "ES1", "ES2", "ES3", "ES4", "E_H", "E_H2", "E_H3", "E_MASK"]
EO, KEEP, STATS = _lib.BLKF_EDGE_OUT, _lib.BLKF_KEEP, _lib.BLKF_STATS


def programs():
    return {
        "fwd[edge_out,keep,stats]": run("dg_block_fwd", FWD_NODE + ["Y_OUT", "Y3", "A16", "E", "Z4", "STAT_M", "STAT_INV"], EO | KEEP | STATS, grads=False),
        "fwd[edge_out]": run("dg_block_fwd", FWD_NODE + ["Y_OUT", "Y3", "A16"], EO, grads=False),
        "fwd[no edge output]": run("dg_block_fwd", FWD_NODE + ["E", "Y3"], 0, grads=False),
        "bwd[kept,weight gradients]": run("dg_block_bwd", BWD_COMMON + BWD_LIVE, EO | KEEP | STATS),
        "bwd[recompute,forward stats,dgrad only]": run("dg_block_bwd", BWD_COMMON + BWD_LIVE, EO | STATS, grads=False),
        "bwd[no edge output,weight gradients]": run("dg_block_bwd", BWD_COMMON + ["Y3"], 0, dead_edge_params=True),
        "bwd_bwd[kept]": run("dg_block_bwd_bwd", BB_COMMON + BB_LIVE, EO | KEEP, second_order=True),
        "bwd_bwd[recompute]": run("dg_block_bwd_bwd", BB_COMMON + BB_LIVE, EO, second_order=True),
        "bwd_bwd[no edge output]": run("dg_block_bwd_bwd", BB_COMMON + ["Y3"], 0, dead_edge_params=True, second_order=True),
    }


if __name__ == "__main__":
    progs = programs()
    if "--write" in sys.argv:
        json.dump({"shape": {"B": B, "N": N, "D": D, "H": H, "heads": HEADS}, "programs": progs}, open(GOLDEN, "w"), indent=1)
        print("wrote", GOLDEN)
    else:
        for k, v in progs.items():
            print("==", k, "(%d launches)" % len([l for l in v if l.startswith("dg_")]))
            print("\n".join("   " + l for l in v))
