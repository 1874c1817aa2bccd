"""ctypes loader placeholder (filled in with the C-ABI)."""


def cuda_backend():
    raise RuntimeError("libdruggen_b200.so is not built; run `python -c 'import __graft_entry__ as g; g.build()'`")
