"""Summarise `ncu --set full` reports (one line block per captured launch): python scripts/ncu_summary.py a.ncu-rep [b.ncu-rep ...]
Reads the raw page through `ncu -i ... --page raw --csv`; prints duration, DRAM bytes, pipe/memory utilisation and launch shape."""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %peak"),
    ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "tmem pipe %active"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor (hmma) %active"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__cluster_size", "cluster"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print(f"== {rep}: {r[hdr.index('Kernel Name')][:90]}")
        for key, label in WANT:
            if key in hdr:
                i = hdr.index(key)
                print(f"   {label:28s} {r[i]} {units[i]}")
        rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        def to_bytes(v, u):
            return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        print(f"   {'dram traffic (bytes)':28s} {int(to_bytes(r[rd], units[rd]) + to_bytes(r[wr], units[wr]))}")
