"""TEST INFRASTRUCTURE ONLY.  Writes tests/golden/data_metric.npz from the reference's OWN function texts.

``src/data/utils.py`` and ``src/util/utils.py`` import torch_geometric / rdkit at module level and cannot be imported here, so the
two functions on this path -- ``label2onehot`` (src/data/utils.py:15-23) and ``average_agg_tanimoto`` (src/util/utils.py:566-611) --
are cut out of the reference files with ``ast`` and executed UNMODIFIED in a namespace that holds only numpy and torch (all
either of them uses).  Nothing is copied into the repository: the source text lives only in this process.

    python oracle/make_golden_data.py          (build container only: the GPU box has no /root/reference)
"""
import ast
import os
import sys

import numpy as np
import torch

REF = os.environ.get("DRUGGEN_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import data_oracle as orc  # noqa: E402


def reference_function(rel_path, name):
    src = open(os.path.join(REF, rel_path)).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = {"np": np, "torch": torch}
    exec(compile(ast.Module(body=[node], type_ignores=[]), rel_path, "exec"), ns)
    return ns[name]


def main():
    ref_onehot = reference_function("src/data/utils.py", "label2onehot")
    ref_tanimoto = reference_function("src/util/utils.py", "average_agg_tanimoto")
    rng = np.random.default_rng(0)
    out = {}
    # label2onehot on a dense bond-label batch (what to_dense_adj hands it) and on atom labels
    x, ei, ea, batch = orc.synthetic_pyg_batch(6, 9, seed=3)
    adj = orc.to_dense_adj(ei, batch, ea, max_num_nodes=9, batch_size=6)
    out.update(pyg_x=x, pyg_edge_index=ei, pyg_edge_attr=ea, pyg_batch=batch, adj_labels=adj,
               a_tensor=ref_onehot(torch.from_numpy(adj), 5).numpy())
    # Tanimoto: 1024-bit fingerprints, ~5 % density, a few empty rows (0/0 -> 1), a duplicated row (similarity 1)
    stock = (rng.random((300, 1024)) < 0.05).astype(np.uint8)
    gen = (rng.random((130, 1024)) < 0.05).astype(np.uint8)
    stock[7] = 0; gen[3] = 0; gen[11] = stock[20]
    out.update(fp_stock=np.packbits(stock, axis=1), fp_gen=np.packbits(gen, axis=1))
    for agg in ("max", "mean"):
        for p in (1, 2):
            out[f"tan_{agg}_p{p}"] = ref_tanimoto(stock, gen, batch_size=128, agg=agg, device="cpu", p=p, intdiv=True)
    out["tan_max_scalar"] = np.float64(ref_tanimoto(stock, gen, batch_size=5000, agg="max", device="cpu"))
    out["tan_self_mean"] = ref_tanimoto(gen, gen, agg="mean", intdiv=True)                # internal_diversity's call (:562)
    dst = os.path.join(os.path.dirname(HERE), "tests", "golden", "data_metric.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
