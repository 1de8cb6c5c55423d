"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list by
kernel: launches, total device time, share of the captured time and (when captured) DRAM bytes moved and the
resulting GB/s. ncu times are cold-cache and serialised: compare SHARES, not absolutes."""
import collections
import csv
import re
import sys

_TIME = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}
_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def short_name(name):
    m = re.search(r"conv_gemm_kernel<\(int\)(\d+), \(int\)(\d+), \(int\)(\d+)(?:, \(bool\)(\d+))?(?:, \(bool\)(\d+))?>", name) or \
        re.search(r"conv_gemm_kernel<(\d+), (\d+), (\d+)(?:, (\d+))?(?:, (\d+))?>", name)
    if m:
        mode, bn, mt, x3, c2 = m.groups()
        return "gp::conv_gemm_kernel<%s,BN=%s,MT=%s%s%s>" % ("WGRAD" if mode == "1" else "FWD", bn, mt,
                                                              ",X3" if x3 == "1" else "", ",PAIR" if c2 == "1" else "")
    return re.sub(r"\(.*", "", name)[:70]


def main(path, top=40):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [set(), 0.0, 0.0])   # launch ids, us, dram bytes
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit, metric, key = row["Metric Unit"], row["Metric Name"], short_name(row["Kernel Name"])
        agg[key][0].add(row["ID"])
        if metric.startswith("gpu__time_duration"):
            agg[key][1] += v * _TIME.get(unit, 1.0)
        elif metric.startswith("dram__bytes"):
            agg[key][2] += v * _BYTES.get(unit, 1.0)
    tot = sum(a[1] for a in agg.values())
    print("total %.1f us over %d launches (device time under ncu: cold cache, serialised)" % (tot, sum(len(a[0]) for a in agg.values())))
    print("%11s %6s %5s %10s %9s  %s" % ("time us", "share", "n", "DRAM MB", "GB/s", "kernel"))
    for k, (ids, us, by) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        gbs = "%9.0f" % (by / us / 1e3) if by and us else "%9s" % "-"
        print("%11.1f %5.1f%% %5d %10.1f %s  %s" % (us, 100 * us / tot, len(ids), by / 1e6, gbs, k))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
