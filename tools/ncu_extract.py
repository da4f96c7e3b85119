#!/usr/bin/env python
"""Turn what `tools/gpu_ncu.sh` brought back in gpurun_out/ into the extracts kept under profiles/.

    python tools/ncu_extract.py [--round r02]

* `ncu_<name>_raw.csv` (ncu --page raw --csv, ~3000 columns) -> `profiles/<round>_ncu_full_<name>.csv`: the columns below
* `ncu_<name>_source.csv` (ncu --page source --csv)        -> `profiles/<round>_ncu_<name>_source_top.csv`: totals + top 40 lines
* `ncu_launches_bench.csv`                                   -> `profiles/<round>_launches_bench.csv` (copied)
* DRAM bytes per launch of the single-target kernel          -> `profiles/traffic.json` entries (same shape/variant replaced)
"""
import argparse
import csv
import json
import os
import re
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = [
    "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "launch__registers_per_thread", "smsp__inst_executed.sum", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.avg.per_second",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
]
NAMES = ("t1_2p23x1024", "t1_2p20x1024", "t1_2p24x512", "mt_262144x2048x100", "mt_forward", "mt_backward",  # gpu_ncu.sh
         "fp_784x10", "fl_512x16")  # gpu_ncu_fl.sh
BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def raw_extract(src, dst):
    rows = list(csv.reader(open(src)))
    header, units, body = rows[0], rows[1], rows[2:]
    cols = [header.index(name) for name in KEEP if name in header]
    with open(dst, "w", newline="") as handle:
        out = csv.writer(handle)
        out.writerow([header[c] for c in cols])
        out.writerow([units[c] for c in cols])
        for row in body:
            out.writerow([row[c] for c in cols])
    return header, units, body


def source_extract(src, dst, top=40):
    rows = list(csv.reader(open(src)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    header = rows[start]
    body = []
    for r in rows[start + 1:]:
        if r and r[0] in ("Kernel Name", "Address"): # the next kernel's section
            break
        if len(r) == len(header):
            body.append(r)
    samples = header.index("# Samples")
    stall = [i for i, name in enumerate(header) if name.startswith("stall_") and "Not Issued" not in name]
    keep = [0, 1, samples, header.index("Instructions Executed")] + stall
    total = ["total samples", sum(int(r[samples] or 0) for r in body), "", ""] + \
            [sum(int(r[i] or 0) for r in body) for i in stall]
    body.sort(key=lambda r: -int(r[samples] or 0))
    with open(dst, "w", newline="") as handle:
        out = csv.writer(handle)
        out.writerow(rows[0][:2])
        out.writerow(total)
        out.writerow([header[i] for i in keep])
        for row in body[:top]:
            out.writerow([row[i].strip() if i == 1 else row[i] for i in keep])


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--round", default="r02")
    parser.add_argument("--src", default=os.path.join(ROOT, "gpurun_out"))
    parser.add_argument("--only", default="", help="comma-separated capture names (default: all of NAMES)")
    args = parser.parse_args()
    names = tuple(args.only.split(",")) if args.only else NAMES
    prof = os.path.join(ROOT, "profiles")
    traffic_path = os.path.join(prof, "traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else []
    for name in sorted(os.listdir(args.src)):
        path = os.path.join(args.src, name)
        m = re.fullmatch(r"ncu_(.+)_raw\.csv", name)
        if m and m.group(1) in names:
            dst = os.path.join(prof, f"{args.round}_ncu_full_{m.group(1)}.csv")
            header, units, body = raw_extract(path, dst)
            print("wrote", dst)
            t1 = re.fullmatch(r"t1_2p(\d+)x(\d+)", m.group(1))
            if t1 and body:
                row = dict(zip(header, body[0]))
                unit = dict(zip(header, units))
                total = sum(float(row[k]) * BYTES[unit[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
                kernel = row["Kernel Name"]
                rows_, feats = 1 << int(t1.group(1)), int(t1.group(2))
                entry = {"kernel": kernel, "rows": rows_, "features": feats, "dram_bytes_per_launch": total,
                         "source": f"profiles/{os.path.basename(dst)} (ncu --set full --clock-control none of a short bench.py "
                                   "run, dram__bytes_read.sum + dram__bytes_write.sum)"}
                for old in traffic:
                    if old.get("rows") == rows_ and old.get("features") == feats and "round 1" not in old.get("source", ""):
                        entry["variant"] = old.get("variant")
                        old.update(entry)
                        break
                else:
                    traffic.append(entry)
        m = re.fullmatch(r"ncu_(.+)_source\.csv", name)
        if m and m.group(1) in names:
            dst = os.path.join(prof, f"{args.round}_ncu_{m.group(1)}_source_top.csv")
            source_extract(path, dst)
            print("wrote", dst)
        if name == "ncu_launches_bench.csv" and not args.only:
            shutil.copy(path, os.path.join(prof, f"{args.round}_launches_bench.csv"))
    json.dump(traffic, open(traffic_path, "w"), indent=1)
    print("updated", traffic_path)


if __name__ == "__main__":
    main()
