#!/usr/bin/env python3
"""ncu CSV launch list of tools/one_step.py -> per-kernel totals of the LAST step: launches, time, DRAM bytes (per step and per cell).
    python tools/dram_bytes.py gpurun_out/launches_su.csv su 2196 [gpurun_out/launches_ss.csv ss 10980] > profiles/r02_dram_bytes.json"""
import csv, json, re, sys

KEYS = [("k_coeff", "k_coeff"), ("k_small", "k_small"), ("k_task_prep", "k_small"), ("k_gram_sum", "k_gram_sum_eval"), ("k_gram_eval", "k_gram_sum_eval"), ("k_gram_interp", "k_gram_sum_eval"),
        ("k_gram", "k_gram"), ("k_contract", "k_contract"), ("k_finalize", "k_finalize"), ("k_gsf", "k_gsf"), ("k_psd", "k_psd"),
        ("k_bessel", "k_bessel"), ("k_pt_table", "k_pt_table"), ("k_xinv", "k_xinv")]
NAMES = {"su": "optics_SU", "ss": "optics_SS"}


def parse(path):
    rows = []
    with open(path) as fp:
        lines = [l for l in fp if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ix = {h: i for i, h in enumerate(hdr)}
    cur = {}
    for r in rd:
        key = r[ix["ID"]]
        name = r[ix["Kernel Name"]]
        metric, val = r[ix["Metric Name"]], float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        if unit in ("Kbyte", "KB"): val *= 1e3
        if unit in ("Mbyte", "MB"): val *= 1e6
        if unit in ("Gbyte", "GB"): val *= 1e9
        if unit in ("us", "usecond"): val *= 1e3
        if unit in ("ms", "msecond"): val *= 1e6
        if unit in ("s", "second"): val *= 1e9
        if key not in cur:
            cur[key] = {"name": name}
            rows.append(cur[key])
        cur[key][metric] = val
    return rows


def kernel_key(name):
    for pat, key in KEYS:
        if pat in name:
            return key
    return None


out = {"how": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none of tools/one_step.py "
              "(launches of the last dense step incl. the GSF stage; per-launch times under ncu are cold-cache and serialised: use shares only)"}
args = sys.argv[1:]
for i in range(0, len(args), 3):
    path, sp, ncell = args[i], args[i + 1], int(args[i + 2])
    rows = parse(path)
    names = [kernel_key(r["name"]) for r in rows]
    # the run ends with two identical steps: the longest suffix that repeats itself is one step
    L = next(n for n in range(len(names) // 2, 0, -1) if names[-n:] == names[-2 * n:-n])
    half = [r for r in rows[-L:] if kernel_key(r["name"]) is not None]
    tot = {}
    for r in half:
        k = kernel_key(r["name"])
        t = tot.setdefault(k, {"launches": 0, "ns": 0.0, "dram_read": 0.0, "dram_write": 0.0})
        t["launches"] += 1
        t["ns"] += r.get("gpu__time_duration.sum", 0.0)
        t["dram_read"] += r.get("dram__bytes_read.sum", 0.0)
        t["dram_write"] += r.get("dram__bytes_write.sum", 0.0)
    total_ns = sum(t["ns"] for t in tot.values())
    d = {"cells": ncell, "launches_in_step": len(half), "step_dram_bytes": sum(t["dram_read"] + t["dram_write"] for t in tot.values()),
         "per_kernel": {k: dict(t, share_of_gpu_time=t["ns"] / total_ns) for k, t in tot.items()}}
    for k, t in tot.items():
        d[k] = (t["dram_read"] + t["dram_write"]) / ncell          # bytes per cell (what bench.py reads)
    out[NAMES[sp]] = d
print(json.dumps(out, indent=1))
