"""Summarise an .ncu-rep here (no GPU needed): python tools/ncu_summary.py file.ncu-rep [kernel-substr] [--src N]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else ""
nsrc = int(sys.argv[sys.argv.index("--src") + 1]) if "--src" in sys.argv else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__waves_per_multiprocessor", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]
for r in rows[2:]:
    name = r[h.index("Kernel Name")]
    if flt not in name:
        continue
    print("===", name[:90])
    for k in KEYS:
        if k in h:
            print("  %-70s %s %s" % (k, r[h.index(k)], rows[1][h.index(k)]))
    st = []
    for k in h:
        if "pcsamp_warps_issue_stalled" in k and not k.endswith("not_issued"):
            try:
                st.append((float(r[h.index(k)]), k.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
    st.sort(reverse=True)
    tot = sum(v for v, _ in st) or 1
    print("  stalls:", ", ".join("%s %.0f%%" % (k, 100 * v / tot) for v, k in st[:7]))
if nsrc:
    # per CUDA source line (needs -lineinfo + --import-source on): samples and executed instructions
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] +
                         (["-k", "regex:" + flt] if flt else []), capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
    h = rows[hi]
    sc, ic = h.index("# Samples"), h.index("Instructions Executed")
    lines = []
    for r in rows[hi + 1:]:
        if r and r[0].strip().isdigit():
            try:
                lines.append((int(r[sc]), int(r[ic]), int(r[0]), r[1].strip()[:100]))
            except ValueError:
                pass
    tot_s = sum(l[0] for l in lines) or 1
    tot_i = sum(l[1] for l in lines) or 1
    print("total samples %d, warp instructions %d" % (tot_s, tot_i))
    for smp, ins, ln, txt in sorted(lines, reverse=True)[:nsrc]:
        print("  %5.1f%% smp %5.1f%% ins  L%-4d %s" % (100.0 * smp / tot_s, 100.0 * ins / tot_i, ln, txt))
