"""profiles/ncu_traffic.json (per kernel: DRAM read + write bytes, launches, time) from an ncu_summary CSV:
    python tools/ncu_traffic.py profiles/r01_v8_ncu_full_summary.csv > profiles/ncu_traffic.json"""
import csv
import json
import re
import sys


def main():
    path = sys.argv[1]
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    col = {h.split(" [")[0]: i for i, h in enumerate(hdr)}
    unit = {h.split(" [")[0]: (h.split("[")[1].rstrip("]") if "[" in h else "") for h in hdr}
    scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
    out = {}
    for r in rows[1:]:
        name = re.sub(r"<.*", "", r[0]).replace("void ", "").split("::")[-1].strip()
        if not r[0].replace("void ", "").startswith("dvcp::"):
            continue
        rd = float(r[col["dram__bytes_read.sum"]]) * scale[unit["dram__bytes_read.sum"]]
        wr = float(r[col["dram__bytes_write.sum"]]) * scale[unit["dram__bytes_write.sum"]]
        k = out.setdefault(name, {"dram_bytes": 0.0, "launches": 0, "time_ms": 0.0, "warp_instructions": 0.0,
                                  "issue_active_pct": 0.0, "tensor_pipe_active_pct": 0.0})
        k["dram_bytes"] += rd + wr
        k["launches"] += 1
        k["time_ms"] += float(r[col["gpu__time_duration.sum"]])
        if "smsp__inst_executed.sum" in col:
            k["warp_instructions"] += float(r[col["smsp__inst_executed.sum"]])
        if "smsp__issue_active.avg.pct_of_peak_sustained_active" in col:
            k["issue_active_pct"] = float(r[col["smsp__issue_active.avg.pct_of_peak_sustained_active"]])
        if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" in col:
            k["tensor_pipe_active_pct"] = float(r[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]])
    print(json.dumps({"source": "%s (ncu --set full --clock-control none, one K8 forward + pose solve, B=8)" % path,
                      "kernels": out}, indent=1))


if __name__ == "__main__":
    main()
