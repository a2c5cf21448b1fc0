"""Extract the headline metrics of every kernel in an .ncu-rep (ncu --set full) into a small text table."""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex_%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_%"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(out.splitlines()))
    hdr, units = rd[0], rd[1]
    print(f"# {path}")
    for row in rd[2:]:
        name = row[hdr.index("Kernel Name")]
        vals = []
        for k, short in KEYS:
            if k in hdr:
                i = hdr.index(k)
                vals.append(f"{short}={row[i]}{units[i] if units[i] not in ('%',) else '%'}")
        print(name[:70], "|", " ".join(vals))


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
