#!/usr/bin/env python
"""tools/ncu_summary.py <report.ncu-rep> [--json-key fwd=regex ...]  -> markdown summary on stdout.

Reads an Nsight Compute report (here, no GPU needed: `ncu -i ... --page raw --csv`) and prints, per
captured launch, the counters the roofline discussion uses.  `--traffic out.json name=regex ...`
additionally writes dram read+write bytes per launch of the first kernel matching each regex
(bench.py reads profiles/roofline_traffic.json for `roofline.traffic`)."""
import csv, io, json, re, subprocess, sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("lts__t_sectors_srcunit_tex_op_red.sum", "L2 reduction sector-ops"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "L1 global-load sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "L1 global-load requests"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid size"),
    ("launch__block_size", "block size"),
    ("smsp__pcsamp_warps_issue_stalled_long_scoreboard", "stall samples: long scoreboard"),
    ("smsp__pcsamp_warps_issue_stalled_lg_throttle", "stall samples: LG throttle"),
    ("smsp__pcsamp_warps_issue_stalled_not_selected", "stall samples: not selected"),
    ("smsp__pcsamp_warps_issue_stalled_wait", "stall samples: wait"),
    ("smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "stall samples: math pipe throttle"),
    ("smsp__pcsamp_sample_count", "stall samples: total"),
]


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def main():
    rep = sys.argv[1]
    hdr, units, rows = load(rep)
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {rep.split('/')[-1]}\n")
    print("`ncu --set full --clock-control none --import-source on` (per-launch values; cold-cache, serialised replays)\n")
    for r in rows:
        print(f"## launch {r[idx['ID']]}: `{r[idx['Kernel Name']]}`\n")
        print("| counter | value | unit |\n|---|---:|---|")
        for k, label in KEYS:
            if k in idx and r[idx[k]] != "":
                v = r[idx[k]]
                try:
                    v = f"{float(v):,.2f}".rstrip("0").rstrip(".")
                except ValueError:
                    pass
                print(f"| {label} (`{k}`) | {v} | {units[idx[k]]} |")
        print()
    if "--traffic" in sys.argv:
        i = sys.argv.index("--traffic")
        path = sys.argv[i + 1]
        res = {}
        for spec in sys.argv[i + 2:]:
            name, rx = spec.split("=", 1)
            for r in rows:
                if re.search(rx, r[idx["Kernel Name"]]):
                    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                    rd = float(r[idx["dram__bytes_read.sum"]]) * scale[units[idx["dram__bytes_read.sum"]]]
                    wr = float(r[idx["dram__bytes_write.sum"]]) * scale[units[idx["dram__bytes_write.sum"]]]
                    res[name] = rd + wr
                    break
        res["_source"] = rep.split("/")[-1]
        json.dump(res, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
