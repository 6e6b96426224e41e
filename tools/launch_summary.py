"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel -> the table in profiles/.

    python tools/launch_summary.py gpurun_out/launches.csv "<command that was profiled>" > profiles/r1_launches_bench_config2.txt
"""
import collections
import csv
import math
import sys


def main():
    path, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    rows = list(csv.reader(ln for ln in open(path) if ln.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, n, tot = collections.OrderedDict(), 0, 0.0
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        if math.isnan(v):
            continue
        us = {"ns": v / 1e3, "nsecond": v / 1e3, "us": v, "usecond": v, "ms": v * 1e3, "msecond": v * 1e3}.get(r[ui], v / 1e3)
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += us
        n += 1
        tot += us
    own = sum(t for k, (c, t) in agg.items() if "bimamba::" in k)
    flush = sum(t for k, (c, t) in agg.items() if "FillFunctor<unsigned char>" in k)
    print(f"# {cmd}")
    print("# (B200, config 2: 4-layer backend fwd+bwd+AdamW, batch 64, 201 frames, bf16).  Per-launch times under ncu are")
    print(f"# serialised and cold-cache: compare SHARES, not absolutes.  {n} launches captured, {tot / 1e3:.1f} ms total,")
    print(f"# of which {flush / 1e3:.1f} ms is bench.py's own 256 MB L2 flush (untimed in the bench); shares below exclude it.")
    work = tot - flush
    print(f"# share of this repository's kernels (bimamba::*): {100 * own / work:.1f} %")
    print("#")
    print("  total_us  count    avg_us  share  kernel")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if "FillFunctor<unsigned char>" in k:
            continue
        print(f"{t:10.1f} {c:6d} {t / c:9.2f} {100 * t / work:5.1f}%  {k[:150]}")


if __name__ == "__main__":
    main()
