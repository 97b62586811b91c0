"""How well do H2D and D2H copies of the cfg2 footprint (76.7 MB each way) overlap on this box, as a function of the copy size and of
the dependency pattern the host-buffer call creates (download g waits for upload g + a kernel)?"""
import time
import torch

dev = torch.device("cuda:0")
TOTAL = 76_677_120 // 4
h_in = torch.empty(TOTAL, dtype=torch.float32).pin_memory(); h_out = torch.empty(TOTAL, dtype=torch.float32).pin_memory()
d_in = torch.empty(TOTAL, dtype=torch.float32, device=dev); d_out = torch.empty(TOTAL, dtype=torch.float32, device=dev)
s_up, s_dn, s_k = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
print("asyncEngineCount", torch.cuda.get_device_properties(dev).multi_processor_count, flush=True)


def run(pieces, mode):
    n = TOTAL // pieces
    ev_up = [torch.cuda.Event() for _ in range(pieces)]; ev_k = [torch.cuda.Event() for _ in range(pieces)]
    ts = []
    for rep in range(8):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        if mode == "independent":
            with torch.cuda.stream(s_up):
                for i in range(pieces): d_in[i * n:(i + 1) * n].copy_(h_in[i * n:(i + 1) * n], non_blocking=True)
            with torch.cuda.stream(s_dn):
                for i in range(pieces): h_out[i * n:(i + 1) * n].copy_(d_out[i * n:(i + 1) * n], non_blocking=True)
        else:
            for i in range(pieces):
                with torch.cuda.stream(s_up):
                    d_in[i * n:(i + 1) * n].copy_(h_in[i * n:(i + 1) * n], non_blocking=True); ev_up[i].record(s_up)
                with torch.cuda.stream(s_k):
                    s_k.wait_event(ev_up[i])
                    if mode == "kernel": d_out[i * n:(i + 1) * n].copy_(d_in[i * n:(i + 1) * n])     # a device kernel between the two copies
                    ev_k[i].record(s_k)
                with torch.cuda.stream(s_dn):
                    s_dn.wait_event(ev_k[i])
                    h_out[i * n:(i + 1) * n].copy_(d_out[i * n:(i + 1) * n], non_blocking=True)
        torch.cuda.synchronize(); ts.append(1e3 * (time.perf_counter() - t0))
    ts.sort()
    return ts[0], ts[len(ts) // 2], ts[-1]


for mode in ("independent", "chain", "kernel"):
    for pieces in (1, 2, 4, 6, 8, 12, 16, 32, 64):
        mn, md, mx = run(pieces, mode)
        print(f"{mode:12s} pieces {pieces:3d} ({TOTAL * 4 / pieces / 1e6:6.2f} MB each): min {mn:.3f} median {md:.3f} max {mx:.3f} ms", flush=True)
