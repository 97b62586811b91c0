"""End-to-end host path (uce_edit_host_f32) on cfg2: wall time per call, and the host-side enqueue time of one call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uce_b200.solver import EditSolver
from uce_b200.synthetic import problem
p = problem("cfg2", seed=0)
s = EditSolver(p["K"], 160, "cuda:0")
h_in = [w.pin_memory() for w in p["W"]]
h_out = [torch.empty_like(w).pin_memory() for w in p["W"]]
hC, hG = p["C"].pin_memory(), p["G"].pin_memory()
for _ in range(3):
    s.edit_host(hC, hG, p["scales"], p["n_edit"], p["lamb"], h_in, h_out)
torch.cuda.synchronize()
ts = []
for _ in range(20):
    t0 = time.perf_counter()
    s.edit_host(hC, hG, p["scales"], p["n_edit"], p["lamb"], h_in, h_out)
    ts.append((time.perf_counter() - t0) * 1e3)
ts.sort()
print(f"edit_host: min {ts[0]:.3f} ms, median {ts[len(ts)//2]:.3f} ms, max {ts[-1]:.3f} ms")
