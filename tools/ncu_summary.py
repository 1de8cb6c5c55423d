"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of captured time)."""
import collections
import csv
import re
import sys


def main(path, top=30):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        name = row["Kernel Name"]
        m = re.search(r"conv_gemm_kernel<\(int\)(\d+), \(int\)(\d+), \(int\)(\d+)>", name) or \
            re.search(r"conv_gemm_kernel<(\d+), (\d+), (\d+)>", name)
        short = ("gp::conv_gemm_kernel<MODE=%s,BN=%s,MT=%s>" % m.groups()) if m else re.sub(r"\(.*", "", name)[:80]
        agg[short][0] += 1
        agg[short][1] += v
        tot += v
    print("total %.1f us over %d launches" % (tot, sum(a[0] for a in agg.values())))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%9.1f us %5.1f%% n=%4d  %s" % (t, 100 * t / tot, n, k))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
