"""Context for the end-to-end number: what the PCIe link of this box gives for the cfg2 footprint (76.7 MB each way),
pinned host memory, one direction at a time and both at once on two streams."""
import time, torch
n = 24960 * 768
h_in, h_out = torch.randn(n).pin_memory(), torch.empty(n).pin_memory()
d_in, d_out = torch.empty(n, device="cuda"), torch.randn(n, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
for _ in range(2): run(True, True, 2)
a, b, c = run(True, False), run(False, True), run(True, True)
print(f"H2D 76.7 MB: {a:.3f} ms ({4*n/a/1e6:.1f} GB/s)   D2H: {b:.3f} ms ({4*n/b/1e6:.1f} GB/s)   both at once: {c:.3f} ms")
# the same bytes as 32 per-projection copies (the granularity uce_edit_host_f32 sees: one host tensor per projection)
from uce_b200_dims import DIMS  # scripts/uce_b200_dims.py
hs_in = [torch.randn(d * 768).pin_memory() for d in DIMS]
hs_out = [torch.empty(d * 768).pin_memory() for d in DIMS]
offs = [0]
for d in DIMS: offs.append(offs[-1] + d * 768)
def run32(h2d, d2h, reps=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                for i, h in enumerate(hs_in): d_in[offs[i]:offs[i + 1]].copy_(h, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                for i, h in enumerate(hs_out): h.copy_(d_out[offs[i]:offs[i + 1]], non_blocking=True)
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
run32(True, True, 2)
a, b, c = run32(True, False), run32(False, True), run32(True, True)
print(f"as 32 per-projection copies: H2D {a:.3f} ms   D2H {b:.3f} ms   both at once {c:.3f} ms")
