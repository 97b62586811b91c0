"""Context for the apply roofline: device-to-device copy of the cfg2 footprint (76.7 MB read + 76.7 MB written per pass),
rotating over several buffers like bench.py does, timed with CUDA events."""
import torch
n = 24960 * 768
R = 3
src = [torch.randn(n, device="cuda") for _ in range(R)]
dst = [torch.empty(n, device="cuda") for _ in range(R)]
for i in range(6):
    dst[i % R].copy_(src[i % R])
torch.cuda.synchronize()
for reps in (30,):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        dst[i % R].copy_(src[i % R])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"copy 76.7 MB -> 76.7 MB: {ms * 1e3:.1f} us per pass, {2 * 4 * n / ms / 1e6:.0f} GB/s")
