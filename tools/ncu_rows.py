"""one line per launch of an .ncu-rep with the metrics that matter for the HBM-bound stages"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(out)))
h = r[0]
want = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue%"), ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bankconf"),
        ("launch__registers_per_thread", "regs"), ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tc%")]
for row in r[2:]:
    name = row[h.index("Kernel Name")]
    name = name[name.find("conv_tc_kernel"):][:48] if "conv_tc_kernel" in name else name[:48]
    print(name, " ".join(f"{lab}={row[h.index(k)][:8]}" for k, lab in want if k in h))
