"""One training step out of an ncu launch list (gpu__time_duration.sum per launch, --csv):

    python tools/launch_summary.py gpurun_out/launches_cunet8_eager.csv [trimmed.csv]

The window ncu captured usually starts and ends mid-step; a step is cut out as [first stem_im2col launch, next
stem_im2col launch) -- the stem's im2col is the first kernel of every step -- so the shares are those of exactly one
forward + backward + optimizer pass.  Prints the per-kernel and per-(kernel, grid) tables; optionally writes the step's
rows back as a CSV for profiles/."""
import collections
import csv
import re
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[start]
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    body = [r for r in rows[start + 2:] if len(r) > vi]
    marks = [i for i, r in enumerate(body) if "stem_im2col" in r[ki]]
    if len(marks) < 2:
        raise SystemExit("the window holds %d stem_im2col launches: capture at least two steps" % len(marks))
    step = body[marks[0]:marks[1]]
    if len(sys.argv) > 2:
        with open(sys.argv[2], "w", newline="") as f:
            w = csv.writer(f, quoting=csv.QUOTE_ALL)
            w.writerows(rows[start:start + 2])
            w.writerows(step)
    per_k, per_kg = collections.defaultdict(lambda: [0, 0.0]), collections.defaultdict(lambda: [0, 0.0])
    for r in step:
        v = float(r[vi].replace(",", ""))
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("cunet::", "")
        name = re.sub(r"<.*", "", name)
        for d, key in ((per_k, name), (per_kg, (name, r[gi]))):
            d[key][0] += 1
            d[key][1] += v
    tot = sum(v[1] for v in per_k.values())
    unit = 1e3 if tot > 1e6 else 1.0          # ncu prints ns or us depending on version
    print("one step: %d launches, %.1f us summed (serialised, cold-cache launch durations)" % (len(step), tot / unit))
    for name, (c, t) in sorted(per_k.items(), key=lambda x: -x[1][1]):
        print("  %-26s %4d launches %9.1f us %5.1f%%" % (name, c, t / unit, 100 * t / tot))
    print("by grid:")
    for (name, g), (c, t) in sorted(per_kg.items(), key=lambda x: -x[1][1])[:24]:
        print("  %-26s %-14s %4d  %9.1f us  avg %6.1f" % (name, g, c, t / unit, t / c / unit))


if __name__ == "__main__":
    main()
