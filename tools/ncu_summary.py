"""Compact summary of `ncu --set full` reports for profiles/:
    python tools/ncu_summary.py out.csv report1.ncu-rep [report2.ncu-rep ...]
One CSV row per report: duration, DRAM bytes / throughput, L2, issue-slot and tensor-pipe utilisation, launch shape."""
import csv, subprocess, sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct_peak"),
    ("dram__bytes.sum.per_second", "dram_bytes_per_s"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct_peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct_peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_active_pct"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "smem_dynamic"),
]


def summarize(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    res = {"report": rep.split("/")[-1], "kernel": d.get("Kernel Name", ("", ""))[1]}
    for k, name in KEYS:
        u, v = d.get(k, ("", ""))
        res[name] = ("%s %s" % (v, u)).strip()
    return res


if __name__ == "__main__":
    dst, reps = sys.argv[1], sys.argv[2:]
    rows = [summarize(r) for r in reps]
    with open(dst, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=list(rows[0].keys()))
        w.writeheader()
        w.writerows(rows)
    for r in rows:
        print(r)
