"""The C-ABI library loads and exports every symbol include/druggen_b200.h declares (no GPU work)."""
import os
import re

import pytest

from druggen_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "druggen_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dg_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip("library not built (run __graft_entry__.build())")
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 16
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(_lib.SIGNATURES) | set(_lib.INFO_SYMBOLS)
    assert lib.dg_abi_version() == _lib.ABI_VERSION


def test_block_slot_table_matches_header():
    """The DG_BLK_* enum of the header (the buffer table of dg_block_fwd / dg_block_bwd / dg_block_bwd_bwd / dg_encoder_fwd) and
    its Python mirror agree slot by slot, as do the flag values and the node-arena size."""
    text = open(os.path.join(ROOT, "include", "druggen_b200.h")).read()
    start = text.index("DG_BLK_X = 0")
    body = text[start:text.index("DG_BLK_COUNT", start)]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = re.findall(r"DG_BLK_([A-Z0-9_]+)", body)
    assert tuple(names) == tuple(_lib.BLK_SLOTS)
    assert _lib.BLK["X"] == 0 and len(set(names)) == len(names)
    for name, val in (("DG_BLKF_EDGE_OUT", _lib.BLKF_EDGE_OUT), ("DG_BLKF_KEEP", _lib.BLKF_KEEP), ("DG_BLKF_STATS", _lib.BLKF_STATS),
                      ("DG_BLOCK_PARAMS", _lib.BLOCK_PARAMS), ("DG_BLK_BB_NODE_SLOTS", _lib.BB_NODE_SLOTS)):
        assert int(re.search(r"#define %s (\d+)" % name, text).group(1)) == val, name
    from druggen_b200.block import BLOCK_PARAM_NAMES
    assert len(BLOCK_PARAM_NAMES) == _lib.BLOCK_PARAMS


def test_header_binds_from_plain_c(tmp_path):
    """include/druggen_b200.h is a C header: examples/block_fwd_trace.c compiles with gcc -std=c99, links against the library and --
    in the dry-run trace, so without a GPU -- lists the ten launches of dg_block_fwd."""
    import shutil
    import subprocess
    if not os.path.exists(_lib.LIB_PATH) or shutil.which("gcc") is None:
        pytest.skip("needs the built library and gcc")
    exe = str(tmp_path / "block_fwd_trace")
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "block_fwd_trace.c"),
                    "-L" + libdir, "-ldruggen_b200", "-Wl,-rpath," + libdir, "-o", exe], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.strip().splitlines()
    launches = [l.split(" ")[0] for l in out[:-1]]
    assert launches == ["dg_add_ln_fwd", "dg_rows_gemm", "dg_rows_gemm", "dg_rows_gemm", "dg_attn_edge_fwd", "dg_softmax_agg16_fwd",
                        "dg_rows_gemm", "dg_add_ln_fwd", "dg_mlp_fwd", "dg_mlp_fwd"]
    assert out[-1].startswith("abi %d, 10 launches" % _lib.ABI_VERSION)
