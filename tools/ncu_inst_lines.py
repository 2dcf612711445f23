"""Executed warp instructions per source line of one kernel from an ncu report captured with --import-source on:
    python tools/ncu_inst_lines.py rep.ncu-rep kernel_regex [top]"""
import csv
import io
import subprocess
import sys


def main():
    rep, kre = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                          "regex:" + kre], capture_output=True, text=True).stdout
    cur_file, out, total = "", [], 0
    col = None
    for r in csv.reader(io.StringIO(raw)):
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            col = {h: i for i, h in enumerate(r)}
            continue
        if r[0] in ("Function Name", "Kernel Name") or col is None or len(r) <= col.get("Instructions Executed", 99):
            continue
        if r[0] and r[2] == "-":      # a source line (not a SASS row)
            try:
                n = int(r[col["Instructions Executed"]] or 0)
            except ValueError:
                continue
            total += n
            out.append((n, cur_file, r[0], r[1].strip()))
    out.sort(reverse=True)
    print("total warp instructions", total)
    for n, f, ln, src in out[:top]:
        print("%10d %5.1f%%  %s:%s  %s" % (n, 100.0 * n / max(total, 1), f, ln, src[:100]))


if __name__ == "__main__":
    main()
