"""The launch programs of the block-level entry points (csrc/block.cu), checked WITHOUT a GPU.

The library's dry-run trace (dg_debug_trace) makes every kernel entry point record its name and arguments instead of
launching; tools/trace_block.py calls dg_block_fwd / dg_block_bwd / dg_block_bwd_bwd with fake, unique addresses per buffer and
maps the recorded pointers back to names.  Two checks:

* the programs equal the committed listing (tests/golden/native_block_programs.json: regenerate with
  ``python tools/trace_block.py --write`` after an intended change and review the diff -- it is the launch list of DESIGN.md 2);
* data flow: a launch only reads buffers that are inputs of the call or were written by an earlier launch, accumulating
  outputs were zeroed (memset0 or the caller's gradient table), inputs and parameters are never written, and every output of
  the call is written.  (Aliasing hazards between scratch slots are what the GPU tests against block.py's launch lists catch.)
"""
import json
import os
import re
import sys

import pytest

from druggen_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.skipif(not os.path.exists(_lib.LIB_PATH), reason="library not built (run __graft_entry__.build())")

# kernel entry -> (positions read, positions written, positions accumulated into = read + written; need a defined value).
# A position may be conditional on an integer argument: (pos, flag_pos, mask) reads the output too when arg[flag_pos] & mask.
SPEC = {
    "dg_add_ln_fwd": ((0, 1, 2, 3), (4,), ()),
    "dg_rows_gemm": ((0, 1, 3, 5, 6), (7,), ()),
    "dg_gemm_tn": ((0, 1), (), (2, 3)),
    "dg_attn_edge_fwd": ((0, 1, 2, 3, 4, 5, 6, 7, 8), (10, 11, 12, 13, 18), ()),
    "dg_softmax_agg16_fwd": ((0, 1), (2, 3, 4), ()),
    "dg_attn_scores_fwd": ((0, 1, 2, 3), (5, 6, 7, 8), ()),
    "dg_attn_scores_bwd": ((0, 1, 2, 3, 4, 5, 7, 8, 9), (10,), (11, 12, 13)),          # de: accumulated into with flag bit 3
    "dg_mlp_fwd": ((0, 1, 2, 3, 4, 5, 6), (7, 12), ()),
    "dg_mlp_bwd_ln": ((0, 1, 2, 3, 4, 5, 6), (7, 8, 9, 16), (10, 11)),
    "dg_mlp_bwd_dgrad": ((0, 1, 2, 3, 4), (5, 6, 10), ()),
    "dg_add_ln_bwd": ((0, 1, 2, 3), (4,), (5, 6)),                                     # dz: accumulated into with `accumulate`
    "dg_add_ln_bwd_bwd": ((0, 1, 2, 3, 4, 5, 6), (7, 8), (9,)),
    "dg_modulate_bwd": ((0, 1, 2, 3), (5, 7), (6,)),
    "dg_modulate_bwd_bwd": ((0, 1, 2, 3, 4, 5, 6), (8, 9, 11), (10,)),
    "dg_softmax_agg_bwd": ((0, 1, 2), (3,), (4,)),                                     # da: accumulated into with `accumulate`
    "dg_softmax_agg_bwd_bwd": ((0, 1, 2, 3, 4), (5, 6), (7,)),
    "memset0": ((), (0,), ()),
    "transpose": ((0,), (1,), ()),
    "add3": ((1, 3, 5), (), (0, 2, 4)),
}
COND_ACC = {"dg_attn_scores_bwd": (10, 17, 8), "dg_add_ln_bwd": (4, 10, 1), "dg_softmax_agg_bwd": (3, 5, 1)}
FWD_KEPT = {"X1", "Q", "K", "V", "ON", "X3", "Y3", "A16", "E", "Z4"}
CASES = {   # program -> (buffers defined on entry besides the parameters / zeroed gradients / scratch vector, outputs)
    "fwd[edge_out,keep,stats]": ({"X", "Y"}, {"X_OUT", "Y_OUT", "Y3", "A16", "E", "Z4", "STAT_M", "STAT_INV", "G", "X1", "Q", "K", "V", "ON", "X3"}),
    "fwd[edge_out]": ({"X", "Y"}, {"X_OUT", "Y_OUT"}),
    "fwd[no edge output]": ({"X", "Y"}, {"X_OUT"}),
    "bwd[kept,weight gradients]": ({"X", "Y", "DXO", "DYO", "STAT_M", "STAT_INV", "G"} | FWD_KEPT, {"DX", "DY"}),
    "bwd[recompute,forward stats,dgrad only]": ({"X", "Y", "DXO", "DYO", "STAT_M", "STAT_INV", "G"}, {"DX", "DY"}),
    "bwd[no edge output,weight gradients]": ({"X", "Y", "DXO"}, {"DX", "DY"}),
    "bwd_bwd[kept]": ({"X", "Y", "DXO", "DYO", "UX", "UY", "X1", "Q", "K", "V", "Y3", "E", "Z4"}, {"C_X", "C_Y", "C_DXO", "C_DYO"}),
    "bwd_bwd[recompute]": ({"X", "Y", "DXO", "DYO", "UX", "UY"}, {"C_X", "C_Y", "C_DXO", "C_DYO"}),
    "bwd_bwd[no edge output]": ({"X", "Y", "DXO", "UX", "UY"}, {"C_X", "C_Y", "C_DXO"}),
}
READ_ONLY = {"X", "Y", "DXO", "DYO", "UX", "UY"}


@pytest.fixture(scope="module")
def programs():
    import trace_block
    return trace_block.programs()


def test_programs_match_the_committed_listing(programs):
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "native_block_programs.json")))["programs"]
    assert set(programs) == set(gold) == set(CASES)
    for name, prog in programs.items():
        assert prog == gold[name], name


def parse(line):
    name, args = re.match(r"(\w+)\((.*)\)$", line).groups()
    return name, [a.strip() for a in args.split(",")]


def base(buf):
    return buf.split("+")[0]


@pytest.mark.parametrize("case", sorted(CASES))
def test_data_flow(programs, case):
    given, outputs = CASES[case]
    defined = set(given) | {"SCRATCH", "WS"}
    kept_inputs = given & FWD_KEPT
    written = set()
    for line in programs[case]:
        name, args = parse(line)
        reads, writes, accs = SPEC[name]
        accs = list(accs)
        if name in COND_ACC:
            pos, flag_pos, mask = COND_ACC[name]
            if int(args[flag_pos]) & mask:
                accs.append(pos)
        ptr = lambda i: args[i] if i < len(args) else "0"  # noqa: E731
        # no launch writes (other than by a declared accumulation) a buffer it also reads: the kernels are not in-place safe
        rd = {ptr(i) for i in reads} - {"0"}
        wr = {ptr(i) for i in writes if i not in accs} - {"0", "WS"}
        assert not (rd & wr), (line, "output aliases an input: %s" % sorted(rd & wr))
        for i in list(reads) + accs:
            b = ptr(i)
            assert not b.startswith("0x"), (line, "unknown address")
            if b == "0" or b.startswith(("P.", "G.")):
                continue
            assert base(b) in defined or b in defined, (line, "reads %s before anything wrote it" % b)
        for i in list(writes) + accs:
            b = ptr(i)
            if b == "0":
                continue
            assert not b.startswith("0x"), (line, "unknown address")
            assert not b.startswith("P."), (line, "writes a parameter")
            assert b not in READ_ONLY and b not in kept_inputs, (line, "writes the call's input %s" % b)
            if b.startswith("G."):
                assert i in accs, (line, "gradient tables are accumulated into, not overwritten")
                continue
            defined.add(b)
            written.add(b)
            if name == "memset0" and b.startswith("N."):          # one fill may cover consecutive node-arena slots
                import trace_block as tb
                k0, slots = tb.ARENA.index(b[2:]), int(args[1]) // (tb.B * tb.N * tb.D * 4)
                defined.update("N." + nm for nm in tb.ARENA[k0:k0 + slots])
    assert outputs <= written, (case, "outputs never written: %s" % sorted(outputs - written))


def test_launch_counts(programs):
    """What DESIGN.md 2 quotes: 10 launches forward, ~25-30 backward, 65-76 for the second-order pass."""
    n = {k: len([l for l in v if l.startswith("dg_")]) for k, v in programs.items()}
    assert n["fwd[edge_out]"] == 10 and n["fwd[no edge output]"] == 9
    assert n["bwd[kept,weight gradients]"] == 24 and n["bwd[recompute,forward stats,dgrad only]"] == 21
    assert n["bwd_bwd[kept]"] == 65 and n["bwd_bwd[recompute]"] == 70
