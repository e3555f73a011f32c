"""Top source lines of an ncu report by warp-stall samples:  python tools/ncu_hot_lines.py report.ncu-rep [N]"""
import csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
cur, rows, tot = None, [], 0
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] not in ("", "Line No", "Function Name") and len(r) > 7:
        try:
            s = int(r[4]); ni = int(r[5]); ex = int(r[7])
        except ValueError:
            continue
        rows.append((s, ni, ex, cur, r[0], r[1].strip()))
        tot += s
rows.sort(reverse=True)
print("total samples", tot)
for s, ni, ex, f, ln, src in rows[:top]:
    print("%6d %5.1f%%  notissued %6d  exec %9d  %s:%s  %s" % (s, 100.0 * s / max(tot, 1), ni, ex, f, ln, src[:110]))
