#!/bin/bash
# ncu launch lists of bench.py itself (final code) and the kernels' shares of the listed time
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench_su.csv \
   python bench.py --workloads su --steps 2 --warmup 3 --no-lut --no-cpu-baseline > gpurun_out/f_ncu_su.log 2>&1; wc -l gpurun_out/r02_launches_bench_su.csv
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02_launches_bench_ss.csv \
   python bench.py --workloads ss --steps 2 --warmup 3 --no-lut --no-cpu-baseline > gpurun_out/f_ncu_ss.log 2>&1; wc -l gpurun_out/r02_launches_bench_ss.csv
python - <<'PY'
import csv, re, collections
out = []
for w in ("ss", "su"):
    rows = [r for r in csv.reader(l for l in open("gpurun_out/r02_launches_bench_%s.csv" % w) if l.startswith('"'))]
    hdr = rows[0]; ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = collections.OrderedDict()
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[ik]).replace("void ", "")
        v = float(r[iv].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "msecond": 1.0, "ms": 1.0}.get(r[iu], 1e-6)
        t = tot.setdefault(name, [0, 0.0]); t[0] += 1; t[1] += v
    s = sum(t[1] for t in tot.values())
    out.append("== bench.py --workloads %s --steps 2 --warmup 3 under ncu (cold-cache, serialised launches): %d launches, %.1f ms listed" % (w, sum(t[0] for t in tot.values()), s))
    for name, t in sorted(tot.items(), key=lambda kv: -kv[1][1])[:12]:
        out.append("   %-34s %6d launches %10.3f ms  %5.1f %%" % (name, t[0], t[1], 100 * t[1] / s))
open("gpurun_out/r02_launch_shares.txt", "w").write("\n".join(out) + "\n")
print("\n".join(out))
PY
