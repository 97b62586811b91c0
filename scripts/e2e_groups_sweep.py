"""Sweep of UCE_HOST_GROUPS for the host-buffer call on the cfg2 footprint, pinned arenas vs one pinned tensor per projection."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uce_b200.solver import EditSolver
from uce_b200.synthetic import SD14_DIMS, problem

dev = torch.device("cuda:0")
K, n, ne, lamb = 768, 150, 50, 0.5
dims = SD14_DIMS
prob = problem("cfg2")
Cr, Gr, scales, W = prob["C"].pin_memory(), prob["G"].pin_memory(), prob["scales"], prob["W"]
_, a_in = EditSolver.host_arena(dims, K); _, a_out = EditSolver.host_arena(dims, K)
for v, w in zip(a_in, W): v.copy_(w)
h_in = [w.pin_memory() for w in W]; h_out = [torch.empty_like(w).pin_memory() for w in W]
for groups in (4, 6, 8, 12, 16, 24, 32):
    os.environ["UCE_HOST_GROUPS"] = str(groups)
    solver = EditSolver(K, n, dev)
    res = []
    for hin, hout in ((a_in, a_out), (h_in, h_out)):
        for _ in range(3): solver.edit_host(Cr, Gr, scales, ne, lamb, hin, hout)
        ts = []
        for _ in range(15):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            solver.edit_host(Cr, Gr, scales, ne, lamb, hin, hout)
            torch.cuda.synchronize(); ts.append(1e3 * (time.perf_counter() - t0))
        ts.sort(); res.append((ts[0], ts[len(ts) // 2]))
    print(f"groups {groups:2d}: arena min {res[0][0]:.3f} median {res[0][1]:.3f} ms | per-tensor min {res[1][0]:.3f} median {res[1][1]:.3f} ms", flush=True)
