"""Per-source-line summary of an ncu report (compile with -lineinfo, capture with --import-source on):
    python tools_ncu_lines.py report.ncu-rep [top_n] [kernel-regex]
Aggregates `ncu --page source --print-source cuda,sass --csv`: for every CUDA source line the executed warp
instructions, the stall samples and the dominant stall reasons."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cmd = ["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"]
if len(sys.argv) > 3:
    cmd += ["-k", "regex:" + sys.argv[3]]
txt = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
fname, hdr, lines = "", None, []
for r in rows:
    if len(r) >= 2 and r[0] in ("File Name", "File Path"):
        fname = r[1].split("/")[-1]
    elif len(r) > 8 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit():
        lines.append((fname, r))
if not lines:
    sys.exit("no source-level data in report")
ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [i for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
tot_i = sum(int(r[ci]) for _, r in lines)
tot_s = sum(int(r[cs]) for _, r in lines)
print(f"total warp instructions {tot_i}, stall samples {tot_s}")
for f, r in sorted(lines, key=lambda t: -int(t[1][cs]))[:top]:
    st = sorted(((int(r[i]), hdr[i][6:]) for i in stall_cols if r[i].isdigit() and int(r[i]) > 0), reverse=True)[:3]
    print(f"{int(r[cs]) / max(tot_s, 1) * 100:5.1f}% smp {int(r[ci]) / max(tot_i, 1) * 100:5.1f}% inst  {f}:{r[0]:>4}  "
          f"{r[1].strip()[:90]:90s} {' '.join(f'{n}={c}' for c, n in st)}")
