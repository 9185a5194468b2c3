"""Turns the ncu exports a gpurun call brings back into the tracked summaries under profiles/:

  python tools/ncu_summarize.py launches gpurun_out/X_launches.csv profiles/Y_ncu_launches_one_step.txt
  python tools/ncu_summarize.py full gpurun_out/X_full_raw.csv profiles/Y_ncu_full_summary.txt [profiles/ncu_traffic.json]

`launches`: per-kernel totals and shares of ONE ELBO step (the launches between two opchain_reset_kernel launches) from
`ncu --metrics gpu__time_duration.sum --clock-control none --csv`.  `full`: the metrics DESIGN.md / bench.py quote from
`ncu --set full` exported with `--page raw --csv`; with a json path, the per-step DRAM traffic of each kernel
(dram__bytes_read.sum + dram__bytes_write.sum summed over its launches of one step) for bench.py's roofline."""
import csv
import json
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"mobo::", "", name)
    return re.sub(r"\(.*$", "", name)


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
    names = [short(r[4]) for r in rows]
    ns = [float(r[-1].replace(",", "")) for r in rows]
    starts = [i for i, n in enumerate(names) if n == "step_prep_kernel"]
    a, b = starts[-2], starts[-1]
    tot, cnt = OrderedDict(), OrderedDict()
    for n, t in zip(names[a:b], ns[a:b]):
        tot[n] = tot.get(n, 0.0) + t / 1e3
        cnt[n] = cnt.get(n, 0) + 1
    total = sum(tot.values())
    with open(dst, "w") as f:
        f.write("# one ELBO step (mobo_elbo_step + mobo_adam) under ncu --metrics gpu__time_duration.sum --clock-control none\n")
        f.write("# bench.py --steps 2 --warmup 3 --no-cpu --no-acq; per-launch times are cold-cache and serialised: compare SHARES\n")
        f.write("# kernel, launches, total us, share\n")
        for n in sorted(tot, key=lambda k: -tot[k]):
            f.write("%-34s %3d %9.1f %5.1f%%\n" % (n, cnt[n], tot[n], 100 * tot[n] / total))
        f.write("%-34s %3d %9.1f\n" % ("TOTAL", b - a, total))


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def full(src, dst, traffic_json=None, note=""):
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    traffic = {}
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none, exported with --page raw --csv (%s)\n" % src)
        f.write("# cold-cache, serialised launches (compare shares, not absolute times)\n\n")
        for r in rows[2:]:
            name = short(r[hdr.index("Kernel Name")])
            f.write("%s  grid %s block %s\n" % (name, r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write("    %-84s %14s %s\n" % (k, r[i], units[i]))
            rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            b = float(r[rd].replace(",", "")) * scale.get(units[rd], 1.0) + float(r[wr].replace(",", "")) * scale.get(units[wr], 1.0)
            traffic.setdefault(name, []).append(b)
    if traffic_json:
        out = {k: v for k, v in traffic.items()}
        json.dump({"per_launch_bytes": out, "_note": note}, open(traffic_json, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None, sys.argv[5] if len(sys.argv) > 5 else "")
