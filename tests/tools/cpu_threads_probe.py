"""How does the CPU restatement of the reference scale with torch threads on this host?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import uce_oracle as O
from uce_b200.synthetic import problem
p = problem("cfg2", seed=0)
ne = p["n_edit"]
W = p["W"][::8]
print("cpu_count", os.cpu_count(), flush=True)
for nt in (2, 4, 8, 12, 16, 32):
    if nt > os.cpu_count():
        continue
    torch.set_num_threads(nt)
    t0 = time.perf_counter()
    O.erase_port_f32(W, p["C"][:ne], p["G"], p["C"][ne:], 1.0, 1.0, 0.5)
    print(f"threads {nt}: {time.perf_counter() - t0:.3f} s for {len(W)} of {len(p['W'])} projections", flush=True)
