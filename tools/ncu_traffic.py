"""DRAM traffic of one training step from an ncu launch list of the bench command:

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
        --log-file L.csv python bench.py --config cfg2 --steps 1 --warmup 3 --no-graph --no-cpu-baseline
    python tools/ncu_traffic.py L.csv cfg2 "<precision label printed by that bench run>" 1024 profiles/r02_....csv

One step = the launches between two consecutive generator optimiser updates (every second gp::adam_flat launch).
Writes / updates profiles/ncu_gemm_traffic.json[<config>], which bench.py reads for `roofline.traffic` — only when its
own configuration, batch and precision label match the capture's."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_TIME = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}
_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def launches(path):
    rows = {}
    order = []
    for row in csv.DictReader(l for l in open(path) if not l.startswith("==")):
        i = int(row["ID"])
        if i not in rows:
            rows[i] = {"name": row["Kernel Name"], "us": 0.0, "bytes": 0.0}
            order.append(i)
        v = float(row["Metric Value"].replace(",", ""))
        if row["Metric Name"].startswith("gpu__time_duration"):
            rows[i]["us"] += v * _TIME.get(row["Metric Unit"], 1.0)
        elif row["Metric Name"].startswith("dram__bytes"):
            rows[i]["bytes"] += v * _BYTES.get(row["Metric Unit"], 1.0)
    return [rows[i] for i in order]


def main(path, config, precision, global_batch, committed_as):
    ls = launches(path)
    adam = [i for i, l in enumerate(ls) if "adam_flat" in l["name"]]
    if len(adam) < 4:
        raise SystemExit("need at least two full steps in the capture (found %d adam_flat launches)" % len(adam))
    # the last COMPLETE step: (generator update k-1, generator update k]
    step = ls[adam[-3] + 1:adam[-1] + 1]
    gemm = [l for l in step if "conv_gemm_kernel" in l["name"]]
    tot_us = sum(l["us"] for l in step)
    out = {
        "dram_bytes_per_step_gemm": sum(l["bytes"] for l in gemm),
        "gemm_launches_per_step": len(gemm),
        "launches_per_step": len(step),
        "gemm_share_of_device_time_under_ncu": sum(l["us"] for l in gemm) / tot_us if tot_us else None,
        "dram_bytes_per_step_all_kernels": sum(l["bytes"] for l in step),
        "source": "%s (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum on `python bench.py "
                  "--config %s --steps 1 --warmup 3 --no-graph --no-cpu-baseline`; one step = the launches between two "
                  "generator optimiser updates; tools/ncu_traffic.py)" % (committed_as, config),
        "precision": precision,
        "global_batch": int(global_batch),
    }
    dst = os.path.join(ROOT, "profiles", "ncu_gemm_traffic.json")
    try:
        allj = json.load(open(dst))
        if "precision" in allj:        # round-1 layout (a single, unkeyed capture)
            allj = {}
    except (OSError, ValueError):
        allj = {}
    allj[config] = out
    json.dump(allj, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(*sys.argv[1:6])
