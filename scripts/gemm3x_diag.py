"""Diagnostic for apply_gemm3x.cu (apply impl 5): error of W_new against the SIMT apply per 32-column block and per 32-row block,
for a few rank pads.  Information for round 2; not a test."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uce_b200.solver import EditSolver
from uce_b200.synthetic import concept_rows, weights

def run(n_edit, K, dims):
    n_pres = 20
    rows = concept_rows(n_edit + n_pres + n_edit, K, seed=n_edit)
    C, G = rows[: n_edit + n_pres].cuda(), rows[n_edit + n_pres:].cuda()
    W = [w.cuda() for w in weights(dims, K, seed=4)]
    s = EditSolver(K, C.shape[0], "cuda:0")
    sc = [1.0] * (n_edit + n_pres)
    s.set_apply_impl(1); a = s.edit(C, G, sc, n_edit, 0.5, W)
    s.set_apply_impl(5); b = s.edit(C, G, sc, n_edit, 0.5, W)
    torch.cuda.synchronize()
    print(f"--- n_edit {n_edit} K {K} dims {dims} info {s.info()}")
    for x, y, w in zip(a, b, W):
        d = (y - x)
        dw = (x - w)
        col = [float(d[:, c:c + 32].norm() / (dw[:, c:c + 32].norm() + 1e-30)) for c in range(0, K, 32)]
        row = [float(d[r:r + 32].norm() / (dw[r:r + 32].norm() + 1e-30)) for r in range(0, x.shape[0], 32)]
        print("  total rel err of the UPDATE", float(d.norm() / dw.norm()))
        print("  per 32-row block:", " ".join(f"{v:.1e}" for v in row))
    s.close()

for d in ([72], [136], [192], [256], [264], [200, 320]):
    run(200, 512, d)
run(40, 512, [200])
