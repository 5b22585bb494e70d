"""Summaries of an `ncu --set full` capture for profiles/: one three-column CSV (metric, unit, value) per captured launch,
headed by the kernel's full name and the hash of the kernel sources it was built from (bench.py reads DRAM traffic and
instruction counts out of these files only when both match the build it is timing).

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/raw.csv
    python tests/tools/ncu_summary.py /tmp/raw.csv profiles/r2_env_kernel_v2_step_ncu_full_summary.csv profiles/r2_env_kernel_v2_reset_ncu_full_summary.csv

Run it from the tree the capture was built from (the hash is taken from the working tree)."""
import csv
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bench import source_hash  # noqa: E402

KEEP = re.compile(r"^(dram__|gpu__time_duration|gpu__dram_throughput|smsp__inst_executed|smsp__thread_inst_executed|smsp__issue_active|sm__warps_active|"
                  r"launch__|sm__inst_executed_pipe_|sm__pipe_(alu|fma|fmaheavy|xu|tensor)_cycles_active|smsp__average_warp|smsp__warps_issue_stalled|"
                  r"l1tex__t_sector_hit_rate|lts__t_sector_hit_rate|sm__throughput|smsp__cycles_active|sm__cycles_elapsed\.max|smsp__pcsamp_warps_issue_stalled)")

raw, outs = sys.argv[1], sys.argv[2:]
rows = list(csv.reader(open(raw)))
hdr, units, launches = rows[0], rows[1], rows[2:]
assert len(launches) >= len(outs), "fewer captured launches than output files"
for launch, out in zip(launches, outs):
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit", "value"])
        w.writerow(["Kernel Name", "", launch[hdr.index("Kernel Name")]])
        w.writerow(["Source Hash", "", source_hash()])
        for name, unit, val in sorted(zip(hdr, units, launch)):
            if KEEP.match(name) and val not in ("", "n/a"):
                w.writerow([name, unit, val])
    print(out, launch[hdr.index("Kernel Name")])
