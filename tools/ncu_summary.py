"""Summarise an .ncu-rep (read on the CPU box): python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.md"""
import csv
import io
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "time",
    "dram__bytes_read.sum": "dram_rd",
    "dram__bytes_write.sum": "dram_wr",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occ_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
}


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [(k, v) for k, v in WANT.items() if k in idx]
    print(f"ncu summary of `{path}` (cold-cache, serialised launches: compare shares, not absolutes)\n")
    print("| kernel | " + " | ".join(f"{v} [{units[idx[k]]}]" for k, v in cols) + " |")
    print("|---|" + "---:|" * len(cols))
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("cruse::", "").replace("<unnamed>::", "")
        print(f"| {name} | " + " | ".join(r[idx[k]] for k, _ in cols) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
