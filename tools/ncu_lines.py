"""Per-source-line stall samples of a kernel from an ncu report captured with
--import-source on (and -lineinfo):  python tools/ncu_lines.py rep.ncu-rep [top]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    cur_file, out, total = "", [], 0
    for r in csv.reader(io.StringIO(raw)):
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] in ("Function Name", "Line No") or len(r) < 6:
            continue
        if r[0] and r[2] == "-":
            try:
                n = int(r[4] or 0)
            except ValueError:
                continue
            total += n
            out.append((n, cur_file, r[0], r[1].strip()))
    out.sort(reverse=True)
    print("total samples", total)
    for n, f, ln, src in out[:top]:
        print("%6d %5.1f%%  %s:%s  %s" % (n, 100.0 * n / max(total, 1), f, ln, src[:110]))


if __name__ == "__main__":
    main()
