"""Builds profiles/r1_summary.md and profiles/r1_traffic.json from the ncu exports kept next to it
(profiles/r1_launches_bench_c2.csv = launch list of the bench command, profiles/raw/*.csv = `ncu -i rep --page raw --csv`
of one `--set full` capture per kernel).  Usage: python profiles/make_summary.py"""
import csv, json, os

HERE = os.path.dirname(os.path.abspath(__file__))
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
CAPTURES = [("fast_pair_warp_kernel (K2+K3 for --bfs gen|sin: warp-autonomous tiles, DMMA contraction), c2 bench step",
             "r1_raw_fast_pair_warp_kernel.csv"),
            ("fast_pair_kernel (K2+K3, CTA-synchronous tile kernel: the --bfs all path; captured on the c2 step with "
             "EQB_FAST_TILE=1 before the warp kernel replaced it there)", "r1_raw_fast_pair_kernel.csv"),
            ("prep_x_dmma_kernel<5,1,16> (K1c, FP64 mma.sync), c2 bench step", "r1_raw_prep_x_dmma.csv"),
            ("prep_y_kernel<12,2> (K1b), c2 bench step", "r1_raw_prep_y_kernel.csv"),
            ("perm_kernel<16,false> (K4, c4 slice of bench.py, --pbf gen-sin; `python profiles/perm_slice_gensin.py`)", "r1_raw_perm_kernel_gensin.csv"),
            ("perm_kernel<8,true> (K4, same slice, --bfs all --pbf all; `python profiles/perm_slice_all.py`)", "r1_raw_perm_kernel_all.csv")]


def raw(path):
    rows = list(csv.reader(open(path)))
    h, units, v = rows[0], rows[1], rows[2]
    return {k: (v[i], units[i]) for i, k in enumerate(h)}


def main():
    out = ["# Round 1 ncu summaries", "",
           "All captures under `gpurun` on one B200 (sm_100a), `--clock-control none`; per-launch times from ncu are "
           "cold-cache and serialised -- compare SHARES. Bench numbers come from `bench.py` (CUDA events), not from these "
           "runs.  Regenerate with `python profiles/make_summary.py` from the exports in `profiles/`.", "",
           "## Launch list of the bench command", "",
           "`ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv python bench.py --steps 2 --warmup 3 "
           "--no-cpu` -> `profiles/r1_launches_bench_c2.csv`", ""]
    rows = [r for r in csv.reader(open(os.path.join(HERE, "r1_launches_bench_c2.csv"))) if len(r) > 10]
    hdr = rows[0]
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    launches = [(r[ki], r[gi], float(r[vi].replace(",", "")) / 1e3) for r in rows[1:]]
    # one timed step = last occurrence of prep_y .. fast_pair before the e2e phase: take the last complete group of the
    # device-only loop (prep_y, prep_x_dmma, prep_x fix-up, fast_pair in a row)
    step = None
    for i in range(len(launches) - 3):
        names = [launches[i + k][0] for k in range(3)]
        if "prep_y_kernel" in names[0] and "prep_x_dmma" in names[1] and "prep_x_kernel" in names[2]:
            j = i + 3
            while j < len(launches) and "stage_copy_kernel" in launches[j][0]:
                j += 1  # work-list uploads of the step (three small staged copies)
            if j < len(launches) and "fast_pair" in launches[j][0]:
                step = launches[i:j + 1]
    if step:
        tot = sum(x[2] for x in step)
        out += ["Kernels of ONE timed step of the c2 workload (251,250 pairs; K1b, K1c, fix-up, work-list uploads, K2+K3):", "",
                "| kernel | grid | us | share |", "|---|---|---|---|"]
        for n, g, us in step:
            out.append("| %s | %s | %.1f | %.1f%% |" % (n.split("(")[0].replace("void ", ""), g, us, 100 * us / tot))
        out += ["| total | | %.1f | 100%% |" % tot, ""]
    agg = {}
    for n, g, us in launches:
        k = n.split("(")[0].replace("void ", "")
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
    out += ["All launches of the command (context set-up, warm-up, timed steps, end-to-end steps, permutation slice):", "",
            "| kernel | launches | total us |", "|---|---|---|"]
    for k, (c, us) in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append("| %s | %d | %.1f |" % (k, c, us))
    out.append("")
    traffic = {}
    for title, fn in CAPTURES:
        p = os.path.join(HERE, "raw", fn)
        if not os.path.exists(p):
            continue
        m = raw(p)
        out += ["## %s -- `ncu --set full --import-source on --clock-control none -k regex:... -c 1`" % title, "",
                "| metric | value |", "|---|---|"]
        for w in WANT:
            if w in m:
                out.append("| %s | %s %s |" % (w, m[w][0], m[w][1]))
        out.append("")

        def tobytes(key):
            v, u = m[key]
            f = float(v.replace(",", ""))
            return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        traffic[fn.replace("r1_raw_", "").replace(".csv", "")] = {
            "dram_bytes_read": tobytes("dram__bytes_read.sum"), "dram_bytes_write": tobytes("dram__bytes_write.sum"),
            "time_us_under_ncu": float(m["gpu__time_duration.sum"][0].replace(",", "")) *
            {"us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6}[m["gpu__time_duration.sum"][1]]}
    open(os.path.join(HERE, "r1_summary.md"), "w").write("\n".join(out) + "\n")
    tj = {"source": "ncu --set full, one capture per kernel (profiles/raw/*.csv)", "kernels": traffic}
    for k, v in traffic.items():
        tj[k + "_dram_bytes_per_launch"] = v["dram_bytes_read"] + v["dram_bytes_write"]
    json.dump(tj, open(os.path.join(HERE, "r1_traffic.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
