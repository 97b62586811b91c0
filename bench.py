#!/usr/bin/env python
"""bench.py — the UCE edit-solve hot path on B200 (BASELINE.json metric: concepts/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2]

A *step* is one complete edit solve of the workload: shared factor + apply over every
cross-attention k/v projection (BASELINE configs[1] by default: SD-1.4 shapes, 50 erase +
100 preserve concepts, 32 projections, K=768).  `value` = concepts/s with the concept rows and
projection weights already resident in HBM (CUDA-graph replay, CUDA-event timed, max over
ranks); `e2e` = the same metric through the host-buffer C-ABI call (uce_edit_host_f32) with
W_old / W_new each in one pinned host arena, H2D and D2H inside the timed region (the same call on one
pinned tensor per projection is reported beside it).  N>1: every rank solves its own
independent edit job (weak scaling, no data-path collective); the layer-sharded single job with
its all-gather is reported under "sharded".  `--impl reference` times the CPU restatement of the
reference's algorithm (oracle/uce_oracle.py: the reference itself is Python and cannot travel
to the GPU box) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

_T0 = time.time()
# stdout carries exactly ONE JSON line: anything a library prints on fd 1 (NCCL's version banner, for one) goes to stderr instead
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def log(msg):
    print(f"[bench +{time.time() - _T0:6.1f}s] {msg}", file=sys.stderr, flush=True)


def host_threads():
    """Threads for the CPU arm.  The reference's loop is thousands of tiny rank-1 updates and 768x768 inverses;
    measured on the 128-core GPU-box host (profiles/cpu_threads_probe_r01.txt) torch's intra-op pool is fastest at
    8 threads (4: 1.8x slower, 16: 2.1x, 32: 12x, 64: 87x, 128: 1950x slower), so 8 is "all the threads it can
    use"; override with UCE_CPU_THREADS."""
    return int(os.environ.get("UCE_CPU_THREADS", min(os.cpu_count() or 1, 8)))


METRIC = "concepts/sec (edit-solve)"
UNIT = "concepts/s"
WORKLOAD_DESC = {
    "cfg1": "cfg1: SDv1.4-shaped, erase 2 preserve 3, one attn2 (to_k,to_v [320,768])",
    "cfg2": "cfg2: SDv1.4 erase 50 preserve 100, all 32 cross-attn k/v projections, K=768",
    "cfg3": "cfg3: SDv1.4 debias-shaped 10 edit rows, 32 projections, K=768",
    "cfg4": "cfg4: SDXL erase 1000, 140 projections, K=2048",
}


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------ CPU baseline
def cpu_port_time(prob, budget_s=20.0, min_reps=1, stratify=1):
    """Time the fp32 restatement of the reference (oracle.erase_port_f32) on the host cores."""
    from oracle import uce_oracle as O
    torch.set_num_threads(host_threads())
    ne = prob["n_edit"]
    ce, cp = prob["C"][:ne], prob["C"][ne:]
    W = prob["W"][::stratify]
    frac = sum(w.shape[0] for w in W) / sum(w.shape[0] for w in prob["W"])
    reps, t_tot = 0, 0.0
    while reps < min_reps or (t_tot < budget_s / 2 and reps < 50):
        t0 = time.perf_counter()
        O.erase_port_f32(W, ce, prob["G"], cp, 1.0, 1.0, prob["lamb"])
        t_tot += time.perf_counter() - t0
        reps += 1
    return t_tot / reps, reps, len(W), frac


def run_reference(args, rank, world):
    from uce_b200.synthetic import problem
    if rank != 0:
        return
    prob = problem(args.workload, seed=0)
    n = prob["C"].shape[0]
    t1, _, _, _ = cpu_port_time(prob, budget_s=0.0, min_reps=1)          # warm-up + sizing
    stratify = 1
    if (args.steps + args.warmup) * t1 > 150.0:
        stratify = 4
    for _ in range(max(0, args.warmup - 1)):
        cpu_port_time(prob, 0.0, 1, stratify)
    ts = []
    frac = 1.0
    for _ in range(args.steps):
        t, _, nl, frac = cpu_port_time(prob, 0.0, 1, stratify)
        ts.append(t / frac if stratify > 1 else t)
    ms = 1e3 * sum(ts) / len(ts)
    val = n / (ms / 1e3)
    sample = (f"full workload per step ({len(prob['W'])} projections)" if stratify == 1 else
              f"every {stratify}th projection per step, time scaled by rows fraction {frac:.3f}")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC[args.workload], "implementation": "CPU fp32 restatement of uce_sd_erase.py:45-82 (oracle port)"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": host_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------ ours
def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from uce_b200.solver import EditSolver
    from uce_b200.synthetic import problem

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    prob = problem(args.workload, seed=rank)
    n, ne, K, lamb = prob["C"].shape[0], prob["n_edit"], prob["K"], prob["lamb"]
    dims = prob["dims"]
    w_bytes = 4 * K * sum(dims)
    R = max(2, int(-(-400e6 // (2 * w_bytes))))          # weight sets so that the rotating footprint exceeds L2 (126 MB) several times
    R = min(R, 8)
    Cd, Gd = prob["C"].to(dev), prob["G"].to(dev)
    sets_in = [[w.to(dev) for w in prob["W"]]]
    for _ in range(R - 1):
        sets_in.append([w.clone() for w in sets_in[0]])
    sets_out = [[torch.empty_like(w) for w in s] for s in sets_in]
    solver = EditSolver(K, max(16, n), dev)
    if args.apply_impl is not None:
        solver.set_apply_impl(args.apply_impl)

    def step(i):      # one complete edit: uce_edit_dev_f32 (factor + apply; the library overlaps what does not depend on the factor)
        solver.edit(Cd, Gd, prob["scales"], ne, lamb, sets_in[i % R], sets_out[i % R], check=False)

    sampler = ClockSampler(local_rank)
    sampler.start()
    log(f"inputs resident: {R} weight sets, {w_bytes/1e6:.1f} MB each")
    # ---- warm-up (eager), then one CUDA graph per weight set ----
    for i in range(max(args.warmup, 3)):
        step(i)
    solver.check()
    info = solver.info()
    launches_per_step = info["launches_factor"] + info["launches_apply"]
    log(f"eager warm-up done: {info}")
    graphs = None
    if not args.no_graph:
        graphs = []
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for i in range(R):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    step(i)
                graphs.append(g)
        torch.cuda.current_stream(dev).wait_stream(side)
        for i in range(max(args.warmup, 3)):
            graphs[i % R].replay()
    torch.cuda.synchronize(dev)
    log("graphs captured and replayed" if graphs else "eager mode")

    def barrier():
        if world > 1:
            dist.barrier()

    # ---- timed region: exactly K steps ----
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(); torch.cuda.synchronize(dev)
    ev0.record()
    for i in range(args.steps):
        if graphs:
            graphs[i % R].replay()
        else:
            step(i)
    ev1.record()
    torch.cuda.synchronize(dev); barrier()
    total_ms = ev0.elapsed_time(ev1)
    solver.check()
    log(f"timed region: {total_ms / args.steps:.4f} ms/step")

    # ---- per-kernel durations (profile events inside the library, eager launches) ----
    solver.set_profile(True)
    tf, t1, t2 = [], [], []
    for i in range(min(args.steps, 20)):
        step(i)
        a, b, c = solver.timings()
        tf.append(a); t1.append(b); t2.append(c)
    solver.set_profile(False)
    f_ms, a1_ms, a2_ms = (sum(x) / len(x) for x in (tf, t1, t2))
    log(f"profiled: factor {f_ms:.4f} ms, apply {a1_ms:.4f} + {a2_ms:.4f} ms")

    # ---- end to end through the host-buffer C-ABI call ----
    e2e_ms = float("nan")
    clocks = None
    e2e_sep_ms = float("nan")
    if not args.no_e2e:
        hC, hG = prob["C"].pin_memory(), prob["G"].pin_memory()

        def time_host(h_in, h_out):
            for _ in range(3):
                solver.edit_host(hC, hG, prob["scales"], ne, lamb, h_in, h_out)
            steps = max(3, min(args.steps, 20))
            barrier(); torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            for _ in range(steps):
                solver.edit_host(hC, hG, prob["scales"], ne, lamb, h_in, h_out)
            torch.cuda.synchronize(dev)
            return 1e3 * (time.perf_counter() - t0) / steps

        # the caller's weights in two pinned arenas (EditSolver.host_arena): one copy per pipeline group and direction
        _, a_in = EditSolver.host_arena(dims, K)
        _, a_out = EditSolver.host_arena(dims, K)
        for v, w in zip(a_in, prob["W"]):
            v.copy_(w)
        e2e_ms = time_host(a_in, a_out)
        # the same call on one separately allocated pinned tensor per projection (32 + 32 small copies)
        h_in = [w.pin_memory() for w in prob["W"]]
        h_out = [torch.empty_like(w).pin_memory() for w in prob["W"]]
        e2e_sep_ms = time_host(h_in, h_out)
        worst = max(float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30) for a, b in zip(a_out, h_out))
        if worst > 1e-5:
            raise SystemExit(f"bench: arena and per-tensor host paths disagree (max rel {worst:.2e})")
        log(f"e2e: {e2e_ms:.3f} ms/step (pinned arenas), {e2e_sep_ms:.3f} ms/step (one pinned tensor per projection)")
    clocks = sampler.stop()

    # ---- reductions over ranks ----
    if world > 1:
        t = torch.tensor([total_ms, e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms = t.tolist()

    sharded = None
    if world > 1:
        sharded = bench_sharded(solver, prob, dev, rank, world, args)

    denoise = None
    if not args.no_denoise:
        # prompts are sharded data-parallel over the GPUs with no steady-state collective:
        # every rank runs its own denoise loop; rank 0 reports the slowest rank's per-GPU rate and the aggregate.
        try:
            denoise = bench_denoise(dev, args)
        except Exception as exc:      # the solver line is the contract; report the denoise failure instead of losing it
            denoise = {"error": f"{type(exc).__name__}: {exc}"}
            log(f"denoise bench failed: {denoise['error']}")
        try:                          # BASELINE cfg5: --num_images_per_prompt 8 (16 samples per U-Net call)
            denoise["bs8"] = bench_denoise(dev, args, images=8)
        except Exception as exc:
            denoise["bs8"] = {"error": f"{type(exc).__name__}: {exc}"}
            log(f"denoise bs=8 bench failed: {denoise['bs8']['error']}")
        if world > 1:
            allv = [None] * world
            dist.all_gather_object(allv, denoise.get("value"))
            if rank == 0 and all(v is not None for v in allv):
                denoise["per_gpu_steps_per_s"] = allv
                denoise["value"] = min(allv)
                denoise["aggregate_steps_per_s"] = sum(allv)
                denoise["n_gpus"] = world
    if rank == 0:
        ms_per_step = total_ms / args.steps
        value = world * n / (ms_per_step / 1e3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_bw, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
        ms_per_step = total_ms / args.steps
        # ---- roofline of the dominant HBM kernel.  ALGORITHMIC bytes:
        #   fused one-kernel apply (impl 4) ......... W_old read + W_new write + E + Q                       = 8 sum(d) K + 8 r K
        #   K-split apply (default, impl 7) kernel B . W_old read + W_new write + ks partial P reads + Q     = 8 sum(d) K + 4 ks sum(d) R + 4 r K
        #                                  kernel A . W_old read + ks-th of E per slice + partial P writes   = 4 sum(d) K + 4 sum(d) R ks + 4 r K
        impl = args.apply_impl or 0
        rank_pad = 32 * max(1, -(-info["rank"] // 32))
        ksplit = 1
        if impl in (0, 7) and rank_pad <= 64 and not info["dense"]:
            ksplit = 1
            while ksplit < 8 and K % (64 * ksplit) == 0 and (K // ksplit > 512 or ksplit < 2) and K // (2 * ksplit) >= 128:
                ksplit *= 2
        sum_d = sum(dims)
        two_kernel = a2_ms > 0.0
        if impl in (0, 7) and two_kernel and rank_pad <= 64:
            kern = "apply_w_kernel (K-split apply, kernel B: W_new = W_old + (sum of the slices' partial W_old E^T) Q; kernel A runs beside the factor)"
            alg_bytes = 2 * w_bytes + 4 * ksplit * sum_d * rank_pad + 4 * K * rank_pad
            dom_ms = a2_ms
            prof_file = "r02_apply_w_ncu.txt"
            alg_a = w_bytes + 4 * ksplit * sum_d * rank_pad + 4 * K * rank_pad
        elif impl == 4:
            kern = "apply_tc3_kernel (fused one-kernel apply)"
            alg_bytes = 2 * w_bytes + 2 * 4 * K * max(info["rank"], 1)
            dom_ms = a1_ms + a2_ms
            prof_file = "r01_apply_tc3_ncu.txt"
            alg_a = None
        else:
            kern = {1: "apply_simt_p / apply_simt_w (fp32 SIMT)", 5: "gemm3x_kernel x 2 (two-GEMM tcgen05 apply)"}.get(impl, "gemm3x_kernel x 2 (two-GEMM tcgen05 apply)")
            alg_bytes = 2 * w_bytes + 2 * 4 * sum_d * rank_pad + 2 * 4 * K * rank_pad      # + P written once, read once
            dom_ms = a1_ms + a2_ms
            prof_file = "r02_gemm3x_ncu.txt"
            alg_a = None
        traffic, traffic_src = None, None          # DRAM bytes of one launch from the committed ncu --set full capture of this command
        try:
            vals = {}
            for ln in open(os.path.join(ROOT, "profiles", prof_file)):
                f = ln.split()
                if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    vals[f[0]] = float(f[1]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[f[2]]
            if len(vals) == 2 and args.workload == "cfg2":
                traffic, traffic_src = sum(vals.values()), f"profiles/{prof_file} (dram__bytes_read.sum + dram__bytes_write.sum of one launch; part of the written rows still sits in L2 at kernel end)"
        except Exception:
            pass
        copy_ref = None
        try:      # context only: a flat device-to-device copy of the same footprint, same rotation, same events
            n_el = w_bytes // 4
            cs = [torch.empty(n_el, device=dev) for _ in range(2 * min(R, 3))]
            hh = len(cs) // 2
            for i in range(4):
                cs[hh + i % hh].copy_(cs[i % hh])
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for i in range(20):
                cs[hh + i % hh].copy_(cs[i % hh])
            c1.record(); torch.cuda.synchronize(dev)
            c_ms = c0.elapsed_time(c1) / 20
            copy_ref = {"what": "torch copy_ of one flat fp32 buffer of the same bytes (read once, write once)", "ms": c_ms, "GB/s": 2 * w_bytes / c_ms / 1e6}
            del cs
        except Exception:
            pass
        achieved = alg_bytes / (dom_ms / 1e3) / 1e9
        step_alg = 2 * w_bytes + 2 * 4 * K * max(info["rank"], 1)          # what ONE edit must move at least: W_old in, W_new out, E, Q
        roof = {"bound": "hbm", "kernel": kern, "achieved": achieved,
                "peak": peak_bw, "unit": "GB/s", "frac": achieved / peak_bw, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes": alg_bytes, "kernel_ms": dom_ms, "stage_ms": [a1_ms, a2_ms], "factor_ms": f_ms,
                "timing": "CUDA events recorded by the library around each kernel on the launching stream (profile mode: kernels serial, eager launches)",
                "copy_reference": copy_ref,
                # the whole step against the bytes one edit must move at least (W_old in, W_new out): what the overlap and the factor cost
                "whole_step": {"algorithmic_bytes": step_alg, "ms": ms_per_step, "GB/s": step_alg / (ms_per_step / 1e3) / 1e9,
                               "frac": step_alg / (ms_per_step / 1e3) / 1e9 / peak_bw},
                # The shared factor: Gram + pack (many CTAs), Cholesky (ONE CTA: 160 serial fp64 pivots), block inverses, solve + emit Q
                # (K / 8 CTAs).  < 1 MB and 40 MFLOP: neither HBM nor a tensor pipe bounds it, latency does.
                "factor": {"kernels": "factor_tables, gram_pack, chol_small (1 CTA), inv_blocks, solve_emit_dmma (chained by programmatic dependent launch)" if info["launches_factor"] <= 5 else "general blocked path (factor.cu)",
                           "ms": f_ms, "bound": "latency (serial fp64 pivots on one SM; dependent fp64 operations cost ~35 cycles each)"}}
        if kern.startswith("gemm3x") and not info["dense"]:
            # the two-GEMM apply at high rank is bound by the tensor pipe, not by HBM: 3 TF32 MMAs (hi.hi, lo.hi, hi.lo) per product,
            # two passes (P = W_old E^T, W_new = W_old + P Q): 12 sum(d) K r_pad flop; kind::tf32 runs at half the bf16 rate
            tf32_peak = 0.5 * peaks.get("bf16_tflops", 2250.0)
            flops = 12.0 * sum_d * K * rank_pad
            tfl = flops / (dom_ms / 1e3) / 1e12
            roof.update({"bound": "tensor", "achieved": tfl, "peak": tf32_peak, "unit": "TFLOP/s", "frac": tfl / tf32_peak,
                         "peak_source": "half of MEASURED_PEAKS.json bf16_tflops (burst): kind::tf32 issues at half the bf16 rate" if "bf16_tflops" in peaks
                         else "half of the nominal 2 250 TFLOP/s bf16 peak", "algorithmic_flops": flops,
                         "hbm": {"achieved_GBs": achieved, "frac": achieved / peak_bw, "algorithmic_bytes": alg_bytes}})
        if alg_a is not None:
            roof["kernel_a"] = {"kernel": "apply_p_kernel (partial W_old E^T per K slice; runs on the library's side stream while the factor computes Q)",
                                "algorithmic_bytes": alg_a, "kernel_ms": a1_ms, "achieved": alg_a / (a1_ms / 1e3) / 1e9, "frac": alg_a / (a1_ms / 1e3) / 1e9 / peak_bw}
            roof["apply_both_kernels"] = {"algorithmic_bytes": step_alg, "ms": a1_ms + a2_ms, "frac": step_alg / ((a1_ms + a2_ms) / 1e3) / 1e9 / peak_bw,
                                          "note": "kernel A + kernel B back to back against the bytes of the fused formulation (the fused one-kernel form, --apply-impl 4, measures 0.47)"}
        gpu_torch = None
        if not args.no_cpu:
            try:      # context: the reference's own execution path — the same fp32 port with torch library kernels on THIS GPU (uce_sd_erase.py runs on cuda:0)
                from oracle import uce_oracle as O
                ce, cp_ = prob["C"][:ne].to(dev), prob["C"][ne:].to(dev)
                Wd = [w.to(dev) for w in prob["W"]]
                O.erase_port_f32(Wd[:2], ce, prob["G"].to(dev), cp_, 1.0, 1.0, lamb)
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                O.erase_port_f32(Wd, ce, prob["G"].to(dev), cp_, 1.0, 1.0, lamb)
                torch.cuda.synchronize(dev)
                t_g = time.perf_counter() - t0
                gpu_torch = {"value": n / t_g, "unit": UNIT, "ms_per_step": 1e3 * t_g,
                             "what": "oracle.erase_port_f32 (the reference's loop order, uce_sd_erase.py:45-82) with torch tensors on this GPU: cuBLAS / cuSOLVER library kernels, one run after a warm-up"}
                log(f"torch-on-GPU port of the reference loop: {1e3 * t_g:.1f} ms per solve")
                del Wd
            except Exception as exc:
                gpu_torch = {"error": f"{type(exc).__name__}: {exc}"}
        cpu = None
        if not args.no_cpu:
            log(f"cpu baseline on {host_threads()} threads ...")
            t_cpu, reps, nl, frac = cpu_port_time(prob, budget_s=args.cpu_budget, min_reps=1)
            log(f"cpu baseline: {t_cpu:.3f} s per solve ({reps} reps)")
            cpu = {"value": n / t_cpu, "unit": UNIT, "cores": host_threads(), "kind": "port",
                   "sample": f"{reps} repetitions of the full {args.workload} workload ({nl} projections), oracle.erase_port_f32, torch {torch.get_num_threads()} threads"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": WORKLOAD_DESC[args.workload], "concept_rows": n, "n_edit": ne, "projections": len(dims), "K": K,
                           "factor": f"fp64 {info['mode_name']} system {info['sys_n']}x{info['sys_n']}, rank {info['rank']}, dense={info['dense']}",
                           "l2": f"inputs rotate over {R} weight sets ({R * 2 * w_bytes / 1e6:.0f} MB in+out) > 126 MB L2, no flush",
                           "launch": "eager" if not graphs else "one CUDA graph replay per step",
                           "parallelism": f"{world} independent edit jobs (one per GPU)" if world > 1 else "1 GPU"},
                "clocks": clocks,
                "e2e": None if args.no_e2e else {"value": world * n / (e2e_ms / 1e3), "unit": UNIT, "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": w_bytes + 4 * K * (n + ne), "d2h_bytes_per_step": w_bytes,
                        "api": "uce_edit_host_f32, W_old / W_new each in one pinned host arena (EditSolver.host_arena)",
                        "ms_per_step_separate_tensors": e2e_sep_ms},
                "gpu_launches": launches_per_step * args.steps,
                "roofline": roof}
        if cpu:
            line["cpu_baseline"] = cpu
        if gpu_torch:
            line["torch_gpu_baseline"] = gpu_torch
        if sharded:
            line["sharded"] = sharded
        if denoise:
            line["denoise"] = denoise
        emit(line)
    solver.close()


# per-op lower bound max(FLOPs / tensor peak, bytes / HBM peak) of one SD-1.4 U-Net call, summed over the ops, at the
# nominal peaks (2.25 PFLOP/s, 8 TB/s): NB = 2: 0.902 ms = 0.685 ms in tensor-bound ops + 0.217 ms in HBM-bound ops; NB = 16 (cfg5):
# 6.75 ms = 5.481 + 1.269.  Against MEASURED peaks each part is rescaled by its own ratio (the op classes keep their bound type:
# the measured ridge, 1414.5 TFLOP/s / 6454.6 GB/s = 219 F/B, is below the nominal 281 and every tensor-bound class sits above both).
UNET_BOUND_NOMINAL_MS = {2: (0.6851, 0.2169), 16: (5.4808, 1.2692)}


def unet_bound_ms(nb, tflops, gbs):
    t, h = UNET_BOUND_NOMINAL_MS[nb]
    return t * 2250.0 / tflops + h * 8000.0 / gbs


def bench_denoise(dev, args, images=1):
    """BASELINE metric part 2: denoise-steps/sec/GPU at 512x512 (64x64 latents) with classifier-free guidance (2 * images samples per
    U-Net call: images = 1 is the north_star target, images = 8 is BASELINE cfg5's --num_images_per_prompt 8), SD-1.4 configuration,
    seeded synthetic weights; one CUDA-graph replay per step (U-Net + guidance + scheduler update)."""
    from uce_b200.generate import Denoiser
    from uce_b200.synthetic import unet_random_state
    from uce_b200.unet import UNetEngine, cfg_step
    from uce_b200.unet_spec import SD14, param_count
    NB = 2 * images
    log(f"denoise: building SD-1.4 U-Net (859.5 M synthetic parameters), {NB} samples per call ...")
    state = unet_random_state(SD14, seed=0)
    eng = UNetEngine(SD14, batch=NB, H=64, W=64, device=dev)
    eng.load_state_dict(state)
    del state
    eng.finalize()
    log(f"denoise: engine ready, {eng.launch_count()} kernels per U-Net call")
    den = Denoiser(eng, images)
    g = torch.Generator().manual_seed(2219)
    den.x.copy_(torch.randn(images, 4, 64, 64, generator=g))
    ctx = torch.randn(NB, 77, 768, generator=g).to(dev)
    c = (55 / 24, -59 / 24, 37 / 24, -9 / 24)

    eng.set_context(ctx)      # the prompt embedding is constant over the steps of a row: its K / V^T projections are per-prompt work

    def step():
        den.x2[:images].copy_(den.x); den.x2[images:].copy_(den.x)
        eng.forward(den.x2, 481.0, None, out=den.eps2)
        cfg_step(den.eps2, 7.5, den.x, den.x, c, 1.0, -0.001, hist=den.hist[:3], eps_out=den.scratch)

    for _ in range(3):
        step()
    torch.cuda.synchronize(dev)
    graph = None
    if not args.no_graph:
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                step()
        torch.cuda.current_stream(dev).wait_stream(side)
        for _ in range(3):
            graph.replay()
    torch.cuda.synchronize(dev)
    K = max(5, min(args.steps, 50))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        graph.replay() if graph else step()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / K
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    # a step sits inside a long loop: the sustained bf16 figure is the denominator (B200_PROFILING.md)
    tf = peaks.get("bf16_tflops_sustained", 1400.0)
    bw = peaks.get("hbm_gbs", 6650.0)
    src = "measured (MEASURED_PEAKS.json: bf16_tflops_sustained, hbm_gbs)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md: 1.4 PFLOP/s sustained, 6.65 TB/s)"
    bound_ms = unet_bound_ms(NB, tf, bw)
    flops = 1.608e12 * NB / 2
    out = {"metric": f"denoise-steps/sec/GPU @512x512 (bs={images}, CFG)", "value": 1e3 / ms, "unit": "steps/s", "ms_per_step": ms, "steps": K,
           "image_steps_per_s": images * 1e3 / ms,
           "config": {"model": "SD-1.4 U-Net shapes, synthetic weights", "params": param_count(SD14), "latents": f"{images}x4x64x64", "unet_batch": NB,
                      "dtype": "bf16 storage, fp32 accumulate", "launch": "one CUDA graph replay per step" if graph else "eager",
                      "kernels_per_step": eng.launch_count() + 3,
                      "context": f"text-context K / V^T projections cached per prompt (sd_unet_set_context, {eng.context_launch_count()} launches, outside the step)"},
           "roofline": {"bound": "mixed (per-op max of tensor and HBM time, summed over the op classes of SURVEY Appendix A)", "achieved_ms": ms,
                        "bound_ms": bound_ms, "frac": bound_ms / ms, "peak_tflops": tf, "peak_gbs": bw, "peak_source": src,
                        "bound_ms_at_nominal_peaks": unet_bound_ms(NB, 2250.0, 8000.0),
                        "algorithmic_flops": flops, "achieved_tflops": flops / 1e12 / (ms / 1e3)}}
    log(f"denoise (bs={images}): {ms:.3f} ms/step = {1e3 / ms:.1f} steps/s, {bound_ms / ms:.3f} of the measured-peak bound ({bound_ms:.3f} ms)")
    eng.close()
    return out


def bench_sharded(solver, prob, dev, rank, world, args):
    """One edit job, projections dealt round-robin to ranks, ONE all-gather of the edited weights."""
    import torch.distributed as dist
    from uce_b200.sharding import GatherPlan
    n, ne, K = prob["C"].shape[0], prob["n_edit"], prob["K"]
    from uce_b200.synthetic import problem
    p0 = problem(args.workload, seed=0)                    # every rank: the same job
    Cd, Gd = p0["C"].to(dev), p0["G"].to(dev)
    dims = p0["dims"]
    from uce_b200.sharding import default_chunks, sharded_edit
    chunks = default_chunks(dims, K, world)
    plan = GatherPlan(dims, K, world, rank, dev, chunks=chunks)
    W = {i: p0["W"][i].to(dev) for i in plan.views_mine()}
    def once():
        return sharded_edit(solver, plan, Cd, Gd, p0["scales"], ne, p0["lamb"], W, check=False)      # W_new = slices of the gather buffer
    for _ in range(3):
        once()
    torch.cuda.synchronize(dev); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 10
    for _ in range(reps):
        once()
    e1.record(); torch.cuda.synchronize(dev)
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    return {"ms_per_job": ms, "value": n / (ms / 1e3), "unit": UNIT, "collective": f"{plan.chunks} in-place all_gather_into_tensor (NCCL) of the packed edited weights (apply kernels write into the gather buffer"
            + ("; a chunk is gathered while the next one is computed)" if plan.chunks > 1 else ")"), "chunks": plan.chunks,
            "nvlink_floor_ms": (world - 1) / world * 4.0 * K * sum(dims) / 770e9 * 1e3,      # bytes every GPU must receive / measured peer bandwidth (B200_PROFILING.md)
            "scaling": "strong"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=list(WORKLOAD_DESC), default="cfg2")
    ap.add_argument("--apply-impl", type=int, default=None, help="0 auto, 1 SIMT, 2 tcgen05 (1 tile/CTA), 3 tcgen05 (2 CTAs/SM), 4 tcgen05 (2 row blocks/CTA, balanced wave), 5 / 6 high-rank tcgen05 (opt-in)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-denoise", action="store_true", help="skip the U-Net denoise-step measurement")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end measurement (ncu launch lists: one apply launch per step)")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the UCE hot path has no CPU fallback (use --impl reference for the CPU baseline)")
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"      # NCCL prints its version banner on stdout; stdout carries exactly one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
