"""Compact CSV summary of an .ncu-rep (key roofline / pipe / memory metrics per launch)."""
import csv
import subprocess
import sys

KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__cycles_active.avg"]
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = [hdr.index(k) for k in KEYS if k in hdr]
labels = sys.argv[3].split(",") if len(sys.argv) > 3 else []
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["label"] + [f"{hdr[i]} [{units[i]}]" for i in idx])
    for n, d in enumerate(data):
        w.writerow([labels[n] if n < len(labels) else ""] + [d[i] for i in idx])
print("wrote", out, len(data), "launches")
