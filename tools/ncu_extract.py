#!/usr/bin/env python
"""Turn an `ncu --set full` report into the committed, machine-readable evidence under profiles/:

    python tools/ncu_extract.py gpurun_out/prof.ncu-rep --kernel step_kernel --out profiles/r2_ncu_stepk.csv \
        --traffic-key cfg2_f32_decayed_k8_n1

* writes `--out`: one row per selected metric (duration, registers, shared memory, occupancy limits, instructions,
  issue rate, pipe utilisation, stall reasons, DRAM bytes, shared-memory wavefronts) of the FIRST launch whose
  kernel name contains `--kernel`;
* with `--traffic-key`, records dram__bytes_read.sum + dram__bytes_write.sum of that launch in
  profiles/ncu_traffic.json, which bench.py reads for `roofline.traffic` (no hard-coded literals).
"""
import argparse
import csv
import io
import json
import os
import subprocess

KEEP = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block",
        "launch__occupancy_limit", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.per_cycle_active", "sm__inst_executed_pipe_", "sm__pipe_fma", "sm__pipe_alu",
        "sm__pipe_fmaheavy", "smsp__average_warps_issue_stalled", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared",
        "sm__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "lts__t_bytes.sum", "sm__throughput", "gpu__dram_throughput"]
UNITS = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--kernel", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--traffic-key")
    args = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    name_col = hdr.index("Kernel Name")
    row = next(r for r in rows[2:] if args.kernel in r[name_col])
    out = []
    for k, u, v in zip(hdr, units, row):
        if any(s in k for s in KEEP) and ".max" not in k and ".min" not in k and "per_second" not in k:
            out.append((k, u, v))
    with open(args.out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit", "value"])
        w.writerows(out)
    print(f"wrote {args.out}: {len(out)} metrics of {row[name_col][:80]}")
    if args.traffic_key:
        def val(metric):
            i = hdr.index(metric)
            return float(row[i].replace(",", "")) * UNITS.get(units[i], 1.0)
        total = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
        path = os.path.join(os.path.dirname(os.path.abspath(args.out)), "ncu_traffic.json")
        db = json.load(open(path)) if os.path.exists(path) else {}
        db[args.traffic_key] = {"bytes": total, "kernel": row[name_col], "source": os.path.relpath(args.out),
                                "duration_us": val("gpu__time_duration.sum") if "gpu__time_duration.sum" in hdr else None}
        json.dump(db, open(path, "w"), indent=1, sort_keys=True)
        print(f"{args.traffic_key}: {total / 1e6:.1f} MB of DRAM traffic per launch -> {path}")


if __name__ == "__main__":
    main()
