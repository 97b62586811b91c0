"""Round-2 profile summaries: gpurun_out/ (scratch) -> profiles/r02_* (committed).   python scripts/summarize_profiles_r02.py"""
import collections, csv, json, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def launch_table(path, out, title, cmd):
    if not os.path.isfile(path):
        return
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        k = re.sub(r"\(.*", "", row["Kernel Name"])[:80]
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(row["Metric Value"].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# {title}\n# command: {cmd}\n# per-launch gpu__time_duration.sum (ncu --clock-control none; cold-cache, serialised: compare shares)\n")
        f.write(f"{'launches':>8} {'avg_us':>10} {'share':>7}  kernel\n")
        for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{n:8d} {t / n / 1e3:10.2f} {100 * t / tot:6.1f}%  {k}\n")
    print("wrote", out)


WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_elapsed.max", "sm__cycles_active.avg", "sm__cycles_active.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__m_l1tex2xbar_write_bytes.sum", "lts__t_bytes.sum", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def raw_metrics(rep, out, title, row_index=0):
    if not os.path.isfile(rep):
        return
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# {title}\n# source: {os.path.basename(rep)} (ncu --set full --clock-control none --import-source on; one launch, caches flushed by ncu)\n")
        for r in rows[2 + row_index: 3 + row_index]:
            f.write(f"kernel: {r[hdr.index('Kernel Name')]}\n")
            for w in WANT:
                if w in hdr:
                    f.write(f"{w:75s} {r[hdr.index(w)]:>16s} {units[hdr.index(w)]}\n")
    print("wrote", out)


launch_table(os.path.join(G, "launches.csv"), os.path.join(P, "r02_solver_launches.txt"), "edit-solve (cfg2) kernel launch list: 5 steps of uce_edit_dev_f32",
             "ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 30 python bench.py --no-cpu --no-denoise --no-e2e --no-graph --steps 6 --warmup 3")
launch_table(os.path.join(G, "launches_cfg4.csv"), os.path.join(P, "r02_solver_launches_cfg4.txt"), "edit-solve at BASELINE cfg4 (SDXL shapes, 1000 concepts, 140 projections): kernel launch list of about one step",
             "ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 220 python bench.py --workload cfg4 --no-cpu --no-denoise --no-e2e --no-graph --steps 3 --warmup 2")
launch_table(os.path.join(G, "launches_unet.csv"), os.path.join(P, "r02_unet_launches.txt"), "one SD-1.4 U-Net call (NB=2, 64x64 latents) kernel launch list",
             "ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off python scripts/unet_profile.py")
raw_metrics(os.path.join(G, "prof_apply_w_kernel.ncu-rep"), os.path.join(P, "r02_apply_w_ncu.txt"),
            "apply_w_kernel — kernel B of the K-split apply (the dominant HBM kernel of an edit: W_new = W_old + P Q), cfg2, 148 CTAs")
raw_metrics(os.path.join(G, "prof_apply_p_kernel.ncu-rep"), os.path.join(P, "r02_apply_p_ncu.txt"),
            "apply_p_kernel — kernel A of the K-split apply (partial W_old E^T per K slice; runs beside the factor), cfg2, 148 CTAs")
raw_metrics(os.path.join(G, "prof_solve_emit_dmma_kernel.ncu-rep"), os.path.join(P, "r02_solve_emit_ncu.txt"),
            "solve_emit_dmma_kernel<8> — X = H^-1 Cp per 8-column slab on the fp64 tensor pipe, edit rows -> Q, Qt, tf32 splits; cfg2, 96 CTAs (cold: under ncu its loads do not overlap inv_blocks)")
raw_metrics(os.path.join(G, "prof_gram_pack_kernel.ncu-rep"), os.path.join(P, "r02_gram_pack_ncu.txt"),
            "gram_pack_kernel — H = Cp Cp^T (fp64 split-K, 180 CTAs) + pack rows / E and its tf32 split; cfg2")
raw_metrics(os.path.join(G, "prof_chol_small_kernel.ncu-rep"), os.path.join(P, "r02_chol_small_ncu.txt"), "chol_small_kernel — single-CTA fp64 Cholesky of the 160 x 160 dual system")
raw_metrics(os.path.join(G, "prof_gemm3x.ncu-rep"), os.path.join(P, "r02_gemm3x_pass1_ncu.txt"),
            "gemm3x_kernel pass 1 (P = W_old E^T) — the two-GEMM tcgen05 apply at BASELINE cfg4 (SDXL shapes, 1000 erase concepts, rank pad 1024, K = 2048), first 96 projections", 0)
raw_metrics(os.path.join(G, "prof_gemm3x.ncu-rep"), os.path.join(P, "r02_gemm3x_pass2_ncu.txt"),
            "gemm3x_kernel pass 2 (W_new = W_old + P Q) — the two-GEMM tcgen05 apply at BASELINE cfg4", 1)


def copy_with_header(src, out, header):
    if os.path.isfile(src):
        open(out, "w").write(header + open(src).read())
        print("wrote", out)


copy_with_header(os.path.join(G, "ab_trace.txt"), os.path.join(P, "r02_apply_ab_timeline.txt"),
                 "# K-split apply — per-role timelines of CTA 0 of kernel A and kernel B (clock64 cycles from the CTA's first event), cfg2, warm run\n"
                 "# produced by UCE_AB_TRACE=<file> (serial launches: UCE_NO_OVERLAP=1); columns: kernel, role, item index, up to three timestamps\n"
                 "#   A w_tma i     : TMA for raw W box of item i = (chunk, block) issued\n"
                 "#   A transform i : box landed | A stage free | tcgen05.st done, stage handed to the MMA warp\n"
                 "#   A mma_a i     : barriers passed | 12 tcgen05.mma + commits issued\n"
                 "#   A p 0 / p 2   : all MMAs complete (drain starts) / kernel entry | setup done | teardown\n"
                 "#   B w_tma i     : TMA store of item i = (unit, block) issued\n"
                 "#   B mma_b i     : barriers passed | 24 tcgen05.mma + commits issued\n"
                 "#   B epilogue i  : accumulator ready | addend box landed | box += accumulator done\n"
                 "#   B p 0 / p 2   : P summed, split and stored to tensor memory / kernel entry | setup done | teardown\n")
copy_with_header(os.path.join(G, "chol_trace.txt"), os.path.join(P, "r02_chol_small_phases.txt"),
                 "# chol_small_kernel — phase boundaries of the single factor CTA (UCE_CHOL_TRACE), cfg2 (n = 150 -> 160)\n"
                 "# columns: index, cycles since kernel start, cycles since the previous boundary\n"
                 "# 1 load | 2 potrf of block 0 | per block step: panel (per-row TRSM), next diagonal block updated, lookahead potrf + trailing update + write-out | last: final block written\n"
                 "# rows >= 32 (index, cycles since kernel start), inside the lookahead phases: 32 + 3 kb: warp 1 done with the trailing update, + 1: with the write-out of L,\n"
                 "#   + 2: with clearing H (last step only); 48 + kb: warp 0's potrf of block kb + 1 done — the potrf is the critical path of every lookahead phase\n")
copy_with_header(os.path.join(G, "sanitize_summary.txt"), os.path.join(P, "r02_sanitizer_summary.txt"),
                 "# compute-sanitizer over every kernel family (scripts/sanitize.sh -> tests/tools/sanitize_target.py: low-latency and general factor, K-split /\n"
                 "# fused / two-GEMM tcgen05 apply, SIMT apply, the host-buffer call, one tiny U-Net call, one tiny VAE decode, both CLIP towers; results checked\n"
                 "# against the oracles under the tool; this is the run AFTER the programmatic-launch factor chain and the fp64 tensor-pipe kernels went in)\n"
                 "# memcheck: 0 errors.  synccheck: 0 errors.  racecheck: 0 hazards in the solver kernels; in the U-Net / VAE runs every report is the\n"
                 "# same one — 'Potential RAW hazard (CUDA barrier operation)' on 8 bytes at window offsets 0x58 and 0x1000058 (the second is the same offset in\n"
                 "# the PEER CTA's shared window) in unet_gemm_pair_kernel: the cluster pair signals mbarriers in the other CTA's shared memory (TMA complete_tx with\n"
                 "# cta_group::2, multicast tcgen05.commit), which racecheck, a per-CTA shared-memory tool, reports against the local try_wait reads.  The engine is\n"
                 "# bit-reproducible run to run (tests/test_unet_gpu.py::test_engine_is_bit_reproducible) and matches its oracle on every tap.\n")
for name in ("bench.json", "bench_cfg1.json", "bench_cfg3.json", "bench_cfg4.json", "bench_ref.json"):
    src = os.path.join(G, name)
    if os.path.isfile(src) and os.path.getsize(src) > 10:
        try:
            d = json.loads(open(src).read().strip().splitlines()[-1])
            json.dump(d, open(os.path.join(P, "r02_" + name), "w"), indent=1)
            print("wrote", "profiles/r02_" + name)
        except Exception as exc:
            print("skip", name, exc)
raw_metrics(os.path.join(G, "prof_solve_general.ncu-rep"), os.path.join(P, "r02_solve_general_ncu.txt"),
            "solve_emit_general_kernel (fp64 SIMT, 8 columns per CTA) at BASELINE cfg4 — shared-memory bound (two loads per fma); replaced by solve_emit_dmma_kernel")
raw_metrics(os.path.join(G, "prof_gram_dmma.ncu-rep"), os.path.join(P, "r02_gram_dmma_ncu.txt"),
            "gram_dmma_kernel — H = Cp Cp^T of the general factor on the fp64 tensor pipe (one CTA per 64 x 64 tile, whole K, no atomics) at BASELINE cfg4 (n = 1000, K = 2048)")
raw_metrics(os.path.join(G, "prof_solve_dmma.ncu-rep"), os.path.join(P, "r02_solve_dmma_ncu.txt"),
            "solve_emit_dmma_kernel (fp64 tensor pipe, mma.sync m8n8k4, 16 columns per CTA, L tiles straight from L2) at BASELINE cfg4")


def concat(srcs, out, header):
    parts = [open(os.path.join(G, n)).read() for n in srcs if os.path.isfile(os.path.join(G, n))]
    if parts:
        open(out, "w").write(header + "\n".join(parts))
        print("wrote", out)


concat(["e2e_sweep.txt", "e2e_trace.txt", "pcie_overlap.txt"], os.path.join(P, "r02_e2e_pipeline.txt"),
       "# host-buffer call (uce_edit_host_f32), cfg2 footprint (76.7 MB each way): UCE_HOST_GROUPS sweep (scripts/e2e_groups_sweep.py), device timeline of a few calls\n"
       "# (UCE_HOST_TRACE=1, scripts/e2e_trace.py: per pipeline group the upload, apply and download intervals in ms since the first upload, and when the HOST\n"
       "# submitted them), and what the link itself does for the same bytes (scripts/pcie_overlap_probe.py)\n")
for name in ("cfg5_n1.json", "cfg5_n2.json"):
    if os.path.isfile(os.path.join(G, name)):
        open(os.path.join(P, "r02_" + name), "w").write(open(os.path.join(G, name)).read())
