"""Summarise an .ncu-rep (raw page) into a small text file for profiles/ (run in the build container)."""
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__cluster_size", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__sass_inst_executed_op_utcmma.sum",
    "smsp__inst_executed_op_tma_ld.sum", "smsp__sass_inst_executed_op_tmem_ldt.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as fh:
        fh.write(f"# ncu --set full summary of {rep} (gpurun, B200, --clock-control none); values per launch\n")
        for row in rows[2:]:
            d = dict(zip(hdr, row))
            du = dict(zip(hdr, units))
            fh.write(f"\n== {d.get('Kernel Name', '?')}  grid {d.get('Grid Size', '?')} block {d.get('Block Size', '?')}\n")
            for k in KEYS:
                if k in d and d[k] != "":
                    fh.write(f"{k:75s} {d[k]} {du.get(k, '')}\n")
            for k in sorted(d):
                m = re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active.ratio", k)
                if m and d[k] and float(d[k]) > 0.05:
                    fh.write(f"stall/{m.group(1):69s} {d[k]}\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
