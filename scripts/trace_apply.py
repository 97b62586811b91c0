"""Timeline of CTA 0 of the tcgen05 apply kernel (UCE_TC_TRACE): one warm launch on the cfg2 workload."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uce_b200.solver import EditSolver
from uce_b200.synthetic import problem
p = problem("cfg2", seed=0)
s = EditSolver(p["K"], 160, "cuda:0")
W = [w.cuda() for w in p["W"]]
out = [torch.empty_like(w) for w in W]
os.environ.pop("UCE_TC_TRACE", None)
for _ in range(3):
    s.factor(p["C"].cuda(), p["G"].cuda(), p["scales"], p["n_edit"], p["lamb"]); s.apply(W, out)
torch.cuda.synchronize()
os.environ["UCE_TC_TRACE"] = "gpurun_out/tc_trace.txt"
os.environ["UCE_CHOL_TRACE"] = "gpurun_out/chol_trace.txt"
s.factor(p["C"].cuda(), p["G"].cuda(), p["scales"], p["n_edit"], p["lamb"]); s.apply(W, out)
torch.cuda.synchronize()
print(open("gpurun_out/tc_trace.txt").read())
print("CHOL phases (index, cycles since start, delta):")
print(open("gpurun_out/chol_trace.txt").read())
