"""Condense an `ncu --set full` report into one CSV row per kernel (first captured launch of each).

    python scripts/ncu_summary.py gpurun_out/full.ncu-rep > profiles/rN_ncu_full_summary.csv
"""
import csv
import io
import subprocess
import sys

COLS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__pipe_tensor_subpipe_dmma_cycles_active.sum", "sm__inst_executed_pipe_tensor_subpipe_dmma.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "launch__waves_per_multiprocessor"]


def main(path):
    raw = subprocess.check_output(["ncu", "-i", path, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    cols = [c for c in COLS if c in ix]
    w = csv.writer(sys.stdout)
    w.writerow(["Kernel Name"] + cols)
    w.writerow([""] + [units[ix[c]] for c in cols])
    seen = set()
    for r in body:
        name = r[ix["Kernel Name"]]
        if name in seen or "elementwise" in name:
            continue
        seen.add(name)
        w.writerow([name] + [r[ix[c]] for c in cols])


if __name__ == "__main__":
    main(sys.argv[1])
