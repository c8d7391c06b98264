"""Per-kernel DRAM traffic / duration / tensor-pipe activity from `ncu --set full` reports -> profiles/r1_traffic.json
python tools/ncu_traffic.py A=gpurun_out/r1_full_A.ncu-rep A1=gpurun_out/r1_full_A1.ncu-rep > profiles/r1_traffic.json
The LAST launch of every kernel in a report is taken (the earlier ones include first-launch effects)."""
import csv
import io
import json
import re
import subprocess
import sys

KEYS = {"dram_bytes_read": "dram__bytes_read.sum", "dram_bytes_write": "dram__bytes_write.sum",
        "duration_us": "gpu__time_duration.sum",
        "tensor_pipe_active_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "alu_pipe_active_pct": "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "warp_instructions": "smsp__inst_executed.sum"}
SCALE = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}


def base_name(n):
    n = re.sub(r"^void\s+", "", n)
    n = re.sub(r"<.*", "", n)
    n = re.sub(r"\(.*", "", n)
    return n.split("::")[-1]


out = {}
for arg in sys.argv[1:]:
    wl, rep = arg.split("=", 1)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units = rows[0], rows[1]
    d = {}
    for r in rows[2:]:
        name = base_name(r[h.index("Kernel Name")])
        e = {}
        for k, m in KEYS.items():
            if m in h:
                v = float(r[h.index(m)].replace(",", "") or 0)
                e[k] = v * SCALE.get(units[h.index(m)], 1.0)
        d[name] = e
    out[wl] = d
json.dump(out, sys.stdout, indent=1)
