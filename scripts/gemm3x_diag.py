"""Diagnostic for apply_gemm3x.cu (apply impl 5): error of W_new against the SIMT apply, for a few rank pads, repeated.  For every
launch that is off it prints WHICH rows and WHICH 32-column boxes — the round-1 runs only had per-32-row-block norms, which left open
whether the intermittent error (profiles/r01_gemm3x_diag_*.txt: one chunk's worth of error in about one row of warp 0's 32) sits in the
A operand (then it covers every column of one CTA's [n0, n0 + BN) range for that row: pass 1, or every column of the row: pass 0) or in
the epilogue (then it is confined to one [128 x 32] box).  Information for round 2; not a test."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from uce_b200.solver import EditSolver
from uce_b200.synthetic import concept_rows, weights

BAD = 1e-4            # relative error of the update above which a row / box counts as wrong (clean launches sit at 5e-6)


def ranges(idx):
    """[3,4,5,9] -> '3-5,9'"""
    out, start, prev = [], None, None
    for i in idx:
        if start is None:
            start = prev = i
        elif i == prev + 1:
            prev = i
        else:
            out.append(f"{start}-{prev}" if prev > start else f"{start}"); start = prev = i
    if start is not None:
        out.append(f"{start}-{prev}" if prev > start else f"{start}")
    return ",".join(out)


def run(n_edit, K, dims, repeats=6, impl=5):
    n_pres = 20
    rows = concept_rows(n_edit + n_pres + n_edit, K, seed=n_edit)
    C, G = rows[: n_edit + n_pres].cuda(), rows[n_edit + n_pres:].cuda()
    W = [w.cuda() for w in weights(dims, K, seed=4)]
    s = EditSolver(K, C.shape[0], "cuda:0")
    sc = [1.0] * (n_edit + n_pres)
    s.set_apply_impl(1); a = [t.clone() for t in s.edit(C, G, sc, n_edit, 0.5, W)]
    s.set_apply_impl(impl)
    s.edit(C, G, sc, n_edit, 0.5, W)
    info = s.info()
    print(f"--- impl {impl} n_edit {n_edit} K {K} dims {dims} rank {info['rank']} (pad {-(-info['rank'] // 32) * 32}) dense {info['dense']} apply launches {info['launches_apply']}")
    n_bad = 0
    for rep in range(repeats):
        b = s.edit(C, G, sc, n_edit, 0.5, W)
        torch.cuda.synchronize()
        for li, (x, y, w) in enumerate(zip(a, b, W)):
            d, dw = (y - x).double(), (x - w).double()
            total = float(d.norm() / dw.norm())
            if total <= BAD:
                continue
            n_bad += 1
            row_err = d.norm(dim=1) / (dw.norm(dim=1) + 1e-300)
            bad_rows = [int(i) for i in torch.nonzero(row_err > BAD).flatten().tolist()]
            print(f"  rep {rep} layer {li} (d = {x.shape[0]}): total rel err of the UPDATE {total:.3e}; wrong rows {ranges(bad_rows)} "
                  f"(row mod 128: {ranges(sorted(set(r % 128 for r in bad_rows)))})")
            for r in bad_rows[:8]:
                box_err = [(c // 32, float(d[r, c:c + 32].norm() / (dw[r, c:c + 32].norm() + 1e-300))) for c in range(0, K, 32)]
                bad_boxes = [bx for bx, e in box_err if e > BAD]
                worst = max(e for _, e in box_err)
                print(f"    row {r}: rel err {float(row_err[r]):.3e}, wrong 32-column boxes {ranges(bad_boxes)} of {K // 32} (worst {worst:.2e})")
    print(f"  {n_bad} wrong projection results in {repeats} launches of {len(dims)} projections" if n_bad else f"  clean: {repeats} launches at <= {BAD:g}")
    s.close()
    return n_bad


if __name__ == "__main__":
    for impl in (5, 6):              # 5: A staged in tensor memory (apply_gemm3x.cu); 6: both operands in shared memory (apply_gemm3x_ss.cu)
        bad = 0
        for d in ([72], [136], [192], [256], [264], [200, 320]):
            bad += run(200, 512, d, impl=impl)
        bad += run(40, 512, [200], impl=impl)
        # every round-1 failure had rank pad 224 (BN = 224, 7 chunks in the second GEMM): does the rank pad matter?
        for n_edit in (90, 120, 180, 250):                   # rank pads 96, 128, 192, 256
            bad += run(n_edit, 512, [200, 320], impl=impl)
        bad += run(300, 768, [320, 640, 1280], impl=impl)
        print(f"impl {impl}: wrong results in total", bad)
