"""Device timeline of the host-buffer call (UCE_HOST_TRACE=1) for a few calls at a given UCE_HOST_GROUPS."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["UCE_HOST_TRACE"] = "1"
import torch
from uce_b200.solver import EditSolver
from uce_b200.synthetic import SD14_DIMS, problem
dev = torch.device("cuda:0"); K, n, ne, lamb = 768, 150, 50, 0.5
prob = problem("cfg2")
Cr, Gr, scales, W = prob["C"].pin_memory(), prob["G"].pin_memory(), prob["scales"], prob["W"]
_, a_in = EditSolver.host_arena(SD14_DIMS, K); _, a_out = EditSolver.host_arena(SD14_DIMS, K)
for v, w in zip(a_in, W): v.copy_(w)
solver = EditSolver(K, n, dev)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    solver.edit_host(Cr, Gr, scales, ne, lamb, a_in, a_out)
    torch.cuda.synchronize(); print(f"call {i}: {1e3 * (time.perf_counter() - t0):.3f} ms (includes the trace's own synchronisation)", file=sys.stderr, flush=True)
