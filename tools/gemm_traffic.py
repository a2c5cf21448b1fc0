"""ncu csv (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum for the GEMM launches of one step)
-> profiles/r02_gemm_dram_traffic.json (stamped with the digest of the kernel sources it was taken from): mean DRAM traffic per launch of the dominant kernel."""
import collections
import csv
import json
import sys


def main(path, out):
    with open(path) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    per = collections.defaultdict(dict)
    for r in csv.DictReader(lines):
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}.get(u, 1)
        per[r["ID"]][r["Metric Name"]] = v * scale
        per[r["ID"]]["name"] = r["Kernel Name"]
    rows = [d for d in per.values() if "gemm_bf16" in d.get("name", "")]
    n = len(rows)
    rd = sum(d.get("dram__bytes_read.sum", 0) for d in rows)
    wr = sum(d.get("dram__bytes_write.sum", 0) for d in rows)
    t = sum(d.get("gpu__time_duration.sum", 0) for d in rows)
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from videollama2_b200 import build as vl2_build
    res = {"kernel": "gemm_bf16_tcgen05_kernel", "launches": n, "source_digest": vl2_build._digest(), "dram_read_bytes_total": rd, "dram_write_bytes_total": wr,
           "traffic_bytes_per_launch": (rd + wr) / max(1, n), "time_s_total_under_ncu": t, "source": path}
    json.dump(res, open(out, "w"), indent=1)
    print(res)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
