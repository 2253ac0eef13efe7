"""Summarise an .ncu-rep (read with `ncu -i`, no GPU needed) into the text committed under profiles/."""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active", "sm__pipe_tensor_subpipe_imma_cycles_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warp_issue_stalled", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    name_col = hdr.index("Kernel Name")
    for r in rows[2:]:
        print(f"== kernel: {r[name_col]}")
        for h, u, v in zip(hdr, units, r):
            if any(h.startswith(k) or k in h for k in KEYS) and v != "":
                print(f"  {h} [{u}] = {v}")
    det = subprocess.run(["ncu", "-i", path, "--page", "details", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(det)))
    print("== details (first kernel)")
    for r in rows[1:]:
        if len(r) > 14 and r[0] == "0" and r[11] in ("GPU Speed Of Light Throughput", "Scheduler Statistics", "Warp State Statistics", "Occupancy", "Launch Statistics", "Memory Workload Analysis", "Compute Workload Analysis"):
            print(f"  {r[11]} | {r[12]} | {r[13]} | {r[14]}")


if __name__ == "__main__":
    main(sys.argv[1])
