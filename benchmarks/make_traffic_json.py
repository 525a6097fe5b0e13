#!/usr/bin/env python
"""profiles/fused_conv_traffic.json from an ncu launch list of bench.py's fused-conv launches.

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        -k regex:lowrank2_fused --csv --log-file gpurun_out/fused_launches.csv \
        python bench.py --steps 1 --warmup 1 --no-msca --no-recon --no-lib --no-cpu
    python benchmarks/make_traffic_json.py gpurun_out/fused_launches.csv [--rank 64 --batch 256]

Takes the LAST 13 launches (one trunk step), sums dram bytes, stamps the file with the sha256 of csrc/lowrank2_tc.cu so that bench.py
only quotes it for the kernel it was measured on, and writes the per-layer table next to it (profiles/r02_fused_conv_per_layer_ncu.csv)."""
import argparse
import csv
import hashlib
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--rank", type=int, default=64)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--layers", type=int, default=13)
    args = ap.parse_args()
    rows = [r for r in csv.reader(open(args.csv)) if len(r) > 5]
    hdr = rows[0]
    ki, mi, vi, ui, ii = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
    launches = {}
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        u = r[ui].lower()
        scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3,
                 "nsecond": 1e-3}.get(u, 1)
        launches.setdefault(int(r[ii]), {"kernel": r[ki][:48]})[r[mi]] = v * scale
    ids = sorted(launches)[-args.layers:]
    total = 0.0
    table = []
    for n, i in enumerate(ids):
        L = launches[i]
        rd, wr, us = L.get("dram__bytes_read.sum", 0.0), L.get("dram__bytes_write.sum", 0.0), L.get("gpu__time_duration.sum", 0.0)
        total += rd + wr
        table.append((n + 1, L["kernel"], us, rd, wr))
    sha = hashlib.sha256(open(os.path.join(ROOT, "convnet-approximater_b200", "csrc", "lowrank2_tc.cu"), "rb").read()).hexdigest()[:16]
    out = {"rank": args.rank, "batch": args.batch, "kernel_sha16": sha, "dram_bytes_per_step": total,
           "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum over the 13 fused-conv launches of one bench.py step "
                     "(profiles/r02_fused_conv_per_layer_ncu.csv)"}
    json.dump(out, open(os.path.join(ROOT, "profiles", "fused_conv_traffic.json"), "w"), indent=1)
    with open(os.path.join(ROOT, "profiles", "r02_fused_conv_per_layer_ncu.csv"), "w") as f:
        f.write("# ncu --clock-control none, one trunk step at batch %d, rank %d (cold-cache, serialised launches); kernel sha16 %s\n"
                % (args.batch, args.rank, sha))
        f.write("layer,kernel,gpu_time_us,dram_read_bytes,dram_write_bytes\n")
        for t in table:
            f.write("%d,%s,%.2f,%.0f,%.0f\n" % t)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
