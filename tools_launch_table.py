"""Per-kernel durations of the LAST bench step in an `ncu --metrics gpu__time_duration.sum --csv` launch list.
    python tools_launch_table.py gpurun_out/launches.csv [--md]
Steps are delimited by codebook_prep_kernel launches (first kernel of every forward)."""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
starts = [i for i, r in enumerate(rows) if r[4].startswith("codebook_prep_kernel")]
lo = starts[-1]
step = rows[lo:]
if len(starts) >= 2 and len(step) < starts[-1] - starts[-2]:     # last step truncated by -c: take the one before
    lo, hi = starts[-2], starts[-1]
    step = rows[lo:hi]
tot = sum(float(r[-1]) for r in step) / 1e3
md = "--md" in sys.argv
if md:
    print("| kernel | us | share |\n|---|---|---|")
for r in step:
    us = float(r[-1]) / 1e3
    name = r[4].split("(")[0] if not r[4].startswith("void") else r[4].split("(")[0].replace("void ", "")
    if "coarse_kernel" in r[4]:
        name = r[4].split("(")[0].replace("void ", "")
    print(f"| `{name}` | {us:.1f} | {us / tot * 100:.1f}% |" if md else f"{name:45s} {us:8.1f} us {us / tot * 100:5.1f}%")
print(f"| **total** | {tot:.1f} | |" if md else f"{'total':45s} {tot:8.1f} us")
