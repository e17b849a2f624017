"""Flatten an .ncu-rep (one captured kernel) into a `metric,unit,value` CSV for profiles/.

    python scripts/ncu_summary.py gpurun_out/<tag>/prof_fwd.ncu-rep profiles/<name>.csv
"""
import csv
import re
import subprocess
import sys

# identification, launch geometry, duration and clocks, pipe utilisation, issue slots, memory traffic, stall reasons
KEEP = re.compile(r"^(ID|Kernel Name|Block Size|Grid Size|launch__|gpu__time_duration\.sum|sm__cycles_elapsed\.avg(\.per_second)?$|"
                  r"sm__cycles_active\.avg$|.*pipe_tensor.*(pct|cycles_active).*|"
                  r"sm__inst_executed_pipe_[a-z0-9_]+\.avg\.pct_of_peak_sustained_active$|smsp__issue_active\.avg\.pct|"
                  r"smsp__inst_executed\.sum$|sm__inst_executed\.sum$|dram__bytes_(read|write)\.sum($|\.per_second)|"
                  r"dram__throughput\.avg\.pct|lts__t_bytes\.sum($|\.per_second)|lts__t_sector_hit_rate\.pct|"
                  r"l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum$|sm__warps_active\.avg\.pct|sm__throughput\.avg\.pct|"
                  r"smsp__average_warps?_issue_stalled_[a-z_]+_per_issue_active|smsp__warps_issue_stalled_[a-z_]+\.avg$|"
                  r"smsp__pcsamp_warps_issue_stalled_[a-z_]+$)")


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit", "value"])
        n = 0
        for h, u, v in zip(hdr, units, vals):
            if KEEP.match(h):
                w.writerow([h, u, v])
                n += 1
    print(f"{out}: {n} of {len(hdr)} metrics of {vals[hdr.index('Kernel Name')]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
