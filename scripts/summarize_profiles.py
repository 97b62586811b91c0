"""Turn the ncu outputs brought back in gpurun_out/ into the text summaries committed under profiles/.

    python scripts/summarize_profiles.py r01
"""
import collections
import csv
import os
import re
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)


def launch_table(path, out, title, cmd):
    if not os.path.isfile(path):
        return
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        k = re.sub(r"\(.*", "", row["Kernel Name"])[:80]
        v = float(row["Metric Value"].replace(",", ""))
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# {title}\n# command: {cmd}\n# per-launch gpu__time_duration.sum (ncu --clock-control none; cold-cache, serialised: compare shares)\n")
        f.write(f"{'launches':>8} {'avg_us':>10} {'share':>7}  kernel\n")
        for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{n:8d} {t / n / 1e3:10.2f} {100 * t / tot:6.1f}%  {k}\n")
    print("wrote", out)


def raw_metrics(rep, out, title, wanted):
    if not os.path.isfile(rep):
        return
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# {title}\n# source: {os.path.basename(rep)} (ncu --set full --clock-control none --import-source on)\n")
        for r in rows[2:3]:
            f.write(f"kernel: {r[hdr.index('Kernel Name')]}\n")
            for w in wanted:
                if w in hdr:
                    f.write(f"{w:75s} {r[hdr.index(w)]:>16s} {units[hdr.index(w)]}\n")
    print("wrote", out)


WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_elapsed.max", "sm__cycles_active.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]

launch_table(os.path.join(G, "launches.csv"), os.path.join(P, f"{tag}_solver_launches.txt"), "edit-solve (cfg2) kernel launch list",
             "ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 3 --warmup 3 --no-graph --no-cpu --no-denoise --no-e2e")
launch_table(os.path.join(G, "launches_unet.csv"), os.path.join(P, f"{tag}_unet_launches.txt"), "one SD-1.4 U-Net call (NB=2, 64x64 latents) kernel launch list",
             "ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off python scripts/unet_profile.py")
WANT += ["l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
raw_metrics(os.path.join(G, "prof_apply_tc3.ncu-rep"), os.path.join(P, f"{tag}_apply_tc3_ncu.txt"),
            "apply_tc3_kernel (the apply of the edit solve: two row blocks per CTA, one planned wave)", WANT)
raw_metrics(os.path.join(G, "prof_apply_tc2.ncu-rep"), os.path.join(P, f"{tag}_apply_tc2_ncu.txt"),
            "apply_tc2_kernel (two co-resident CTAs per SM; apply impl 3)", WANT)
raw_metrics(os.path.join(G, "prof_apply_tc.ncu-rep"), os.path.join(P, f"{tag}_apply_tc_ncu.txt"), "apply_tc_kernel (one 128-row tile per CTA; apply impl 2)", WANT)
raw_metrics(os.path.join(G, "prof_chol_small.ncu-rep"), os.path.join(P, f"{tag}_chol_small_ncu.txt"), "chol_small_kernel (single-CTA factor)", WANT)
raw_metrics(os.path.join(G, "prof_unet_gemm_pair.ncu-rep"), os.path.join(P, f"{tag}_unet_gemm_pair_ncu.txt"),
            "unet_gemm_pair_kernel (third pair-GEMM launch of one SD-1.4 U-Net call: 128 CTAs = 64 pairs, 64x64 level, 320 channels)", WANT)
raw_metrics(os.path.join(G, "prof_unet_attn.ncu-rep"), os.path.join(P, f"{tag}_unet_attn_ncu.txt"),
            "unet_attn_kernel (first launch of one SD-1.4 U-Net call: self-attention over 4096 tokens, 8 heads, head dim 40 padded to 64)", WANT)


def copy_with_header(src, out, header):
    if not os.path.isfile(src):
        return
    with open(out, "w") as f:
        f.write(header)
        f.write(open(src).read())
    print("wrote", out)


copy_with_header(os.path.join(G, "tc_trace.txt"), os.path.join(P, f"{tag}_apply_tc3_timeline.txt"),
                 "# apply_tc3_kernel — per-role timeline of CTA 0 (clock64 cycles from the CTA's first event), cfg2 workload, warm run\n"
                 "# produced by UCE_TC_TRACE=<file> (scripts/trace_apply.py); columns: role, index, then up to three timestamps\n"
                 "#   w_tma k      : TMA pair for raw W chunk k issued | (phase B) stores of unit k issued\n"
                 "#   transform k  : raw chunk k landed | A stage free | tcgen05.st issued\n"
                 "#   e_tma k      : TMA for E tile k issued\n"
                 "#   mma_a k      : - | barriers passed (block 0 issuer) | 12 tcgen05.mma + commits issued      (phase A, 24 chunks)\n"
                 "#   mma_b u      : barriers passed | 24 tcgen05.mma + commits issued                        (phase B, 24 units of 32 W columns)\n"
                 "#   epilogue u   : addend box landed | box += accumulator done, handed to the TMA warp\n"
                 "#   pconv        : P final | P_hi / P_lo written back to tensor memory\n")
copy_with_header(os.path.join(G, "chol_trace.txt"), os.path.join(P, f"{tag}_chol_small_phases.txt"),
                 "# chol_small_kernel — phase boundaries of the single factor CTA (UCE_CHOL_TRACE), cfg2 workload (n = 150 -> 160, 50 right-hand sides)\n"
                 "# columns: index, cycles since kernel start, cycles since the previous boundary\n"
                 "# 1 load | per block kb = 0..4: potrf (one warp), inverse of the diagonal block, panel, trailing update | 21 forward | 22 backward | 23 write Z\n")
copy_with_header(os.path.join(G, "fp64_probe.txt"), os.path.join(P, f"{tag}_fp64_probe.txt"),
                 "# scripts/fp64_probe.cu on one SM of a B200: plain DFMA and tensor-core DMMA.8x8x4 both peak at 64 fp64 FMA/clk/SM\n"
                 "# (so the fp64 tensor path offers nothing to the single-CTA factor; its GEMM-like phases are bound by shared-memory wavefronts and latency)\n")
copy_with_header(os.path.join(G, "copy_ceiling.txt"), os.path.join(P, f"{tag}_copy_reference.txt"),
                 "# scripts/copy_ceiling.py: flat device-to-device copy of the cfg2 footprint, rotating buffers, CUDA events (context for roofline.frac)\n")
if os.path.isfile(os.path.join(G, "pcie_floor.txt")):
    with open(os.path.join(P, f"{tag}_e2e_floor.txt"), "w") as f:
        f.write("# Context for the end-to-end (host-buffer) number: what the PCIe link of the GPU box gives for the cfg2 footprint\n"
                "# (scripts/pcie_floor.py), and uce_edit_host_f32 itself (scripts/e2e_probe.py).  The call takes one pinned host tensor per\n"
                "# projection, so 32 uploads and 32 downloads are the granularity it has to work with.\n")
        f.write(open(os.path.join(G, "pcie_floor.txt")).read())
        if os.path.isfile(os.path.join(G, "e2e_probe.txt")):
            f.write(open(os.path.join(G, "e2e_probe.txt")).read())
    print("wrote", os.path.join(P, f"{tag}_e2e_floor.txt"))
