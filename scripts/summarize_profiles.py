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
             "ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 3 --warmup 3 --no-graph --no-cpu --no-denoise")
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
