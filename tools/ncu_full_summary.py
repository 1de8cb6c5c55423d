"""Key counters of every launch in an `ncu --set full` report (.ncu-rep), as text for profiles/:
    python tools/ncu_full_summary.py gpurun_out/r02_prof_gemm.ncu-rep > profiles/r02_ncu_gemm_full_summary.txt
(reads the report with `ncu -i ... --page raw --csv`; needs no GPU)."""
import csv
import io
import subprocess
import sys

KEYS = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum",
        "smsp__average_warp_latency_issue_stalled_no_instruction.ratio"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units = rows[0], rows[1]
    idx = {}
    for i, h in enumerate(head):
        idx.setdefault(h.split(".", 2)[-1] if h.count(".") >= 2 and h.split(".")[1] in ("TriageCompute",) else h, i)
    for r in rows[2:]:
        print("-----")
        print("  %-72s %s" % ("Kernel Name", r[head.index("Kernel Name")]))
        for k in KEYS:
            i = idx.get(k)
            if i is None:
                cands = [j for j, h in enumerate(head) if h.endswith(k)]
                i = cands[0] if cands else None
            if i is not None and r[i] != "":
                print("  %-72s %s %s" % (k, r[i], units[i]))


if __name__ == "__main__":
    main(sys.argv[1])
