#!/usr/bin/env python
"""Summarise `ncu --set full` raw-page CSVs (tools/ncu_capture.sh) per kernel: duration, tensor-pipe activity, DRAM bytes
read + written, L2 throughput -- a text table for profiles/ and profiles/ncu_traffic.json, which bench.py reads for the
`traffic` field of its roofline entries (key = ABI entry point + integer arguments of the launch).
    python tools/ncu_traffic.py gpurun_out/r2_ncu_*.csv --json profiles/ncu_traffic.json > profiles/r2_ncu_full_summary.txt"""
import csv
import json
import os
import re
import sys

# capture name -> bench.py key of the launch it profiles (entry point : dtype, N, Cin, Cout, D, H, W)
KEYS = {
    "conv_32_32": "sg_conv3d_fprop:0,4,32,32,32,128,128",
    "conv_64_32": "sg_conv3d_fprop:0,4,64,32,32,128,128",
    "conv_32_64": "sg_conv3d_fprop:0,4,32,64,32,128,128",
    "wgrad_32_64": "sg_conv3d_wgrad:0,4,32,64,32,128,128",
    "wgrad_32_32": "sg_conv3d_wgrad:0,4,32,32,32,128,128",
}
# tools/membound_bench.py --reps 1 under ncu (cfg3 top-level shapes): one warm-up Adam launch, then 3 launches per case in
# this order; bench.py key of each case (pointer-flag variants carry a suffix, as in bench.roofline_entries)
MEMBOUND_CASES = ["sg_pw_expand:0,4,32,524288,1", "sg_pw_expand_masked:0,4,32,524288,0", "sg_pw_reduce:0,4,32,524288",
                  "sg_pw_wgrad:0,4,32,524288", "sg_pw_wgrad:0,4,64,524288", "sg_down2:0,0,8,32,32,128,128",
                  "sg_up2+mask:0,0,8,32,16,64,64", "sg_up2:0,0,8,32,16,64,64", "sg_mask_mul:0,134217728",
                  "sg_lincomb:0,134217728", "sg_pixelnorm_fwd:0,4,32,524288,1", "sg_pixelnorm_bwd:0,4,32,524288,1,1",
                  "sg_adam_step:32000000"]
COLS = [("gpu__time_duration.sum", "us", 1.0), ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor %", 1.0),
        ("dram__bytes_read.sum", "rd MB", 1.0), ("dram__bytes_write.sum", "wr MB", 1.0),
        ("dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "dram act %", 1.0), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %", 1.0),
        ("launch__registers_per_thread", "regs", 1.0), ("launch__grid_size", "grid", 1.0)]


def to_float(v, unit):
    v = float(v.replace(",", ""))
    scale = {"Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "byte": 1e-6, "ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0,
             "msecond": 1e3}
    return v * scale.get(unit, 1.0)


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    out_json = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
    args = [a for a in args if a != out_json]
    traffic = {}
    print(f"{'capture':14s} {'kernel':58s} " + " ".join(f"{c[1]:>9s}" for c in COLS))
    for path in args:
        name = re.sub(r".*_ncu_", "", os.path.basename(path)).replace(".csv", "")
        rows = list(csv.reader(open(path)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        last = {}
        body = [r for r in rows[2:] if len(r) == len(hdr)]
        if name == "membound" and len(body) == 1 + 3 * len(MEMBOUND_CASES):
            for ci, key in enumerate(MEMBOUND_CASES):
                r = body[1 + 3 * ci + 2]
                g = lambda col: to_float(r[idx[col]], units[idx[col]])      # noqa: E731
                traffic[key] = {"dram_bytes": (g("dram__bytes_read.sum") + g("dram__bytes_write.sum")) * 1e6,
                                "kernel": re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", ""),
                                "ncu_us": g("gpu__time_duration.sum"), "source": os.path.basename(path)}
        for r in rows[2:]:
            if len(r) != len(hdr):
                continue
            last[re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "").replace("<unnamed>::", "")] = r
        for kname, r in last.items():       # the last launch of every kernel: warmed up
            vals = []
            for col, _, _ in COLS:
                vals.append(to_float(r[idx[col]], units[idx[col]]) if col in idx and r[idx[col]] not in ("", "n/a", "no data") else float("nan"))
            print(f"{name:14s} {kname[:58]:58s} " + " ".join(f"{v:9.1f}" for v in vals))
            if name in KEYS and ("k_conv_tc" in kname or "k_wgrad_tc" in kname):
                traffic[KEYS[name]] = {"dram_bytes": (vals[2] + vals[3]) * 1e6, "kernel": kname, "tensor_pipe_pct": vals[1],
                                       "ncu_us": vals[0], "source": os.path.basename(path)}
    if out_json:
        with open(out_json, "w") as f:
            json.dump(traffic, f, indent=1)


if __name__ == "__main__":
    main()
