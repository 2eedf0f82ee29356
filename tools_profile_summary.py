"""Turn gpurun_out/*.ncu-rep and the ncu launch list into the tracked summaries under profiles/.
    python tools_profile_summary.py <tag> <full.ncu-rep> <launches.csv>
"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

tag, rep, launches = sys.argv[1], sys.argv[2], sys.argv[3]
os.makedirs("profiles", exist_ok=True)

WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
kernels = []
for r in rows[2:]:
    d = {"kernel": r[hdr.index("Kernel Name")]}
    for i, h in enumerate(hdr):
        if h in WANT:
            d[h] = f"{r[i]} {units[i]}".strip()
    kernels.append(d)

def mb(s):
    v, u = s.split()[:2]
    return float(v) * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}[u]

dom = kernels[0]
dram = mb(dom["dram__bytes_read.sum"]) + mb(dom["dram__bytes_write.sum"])
json.dump({"kernel": dom["kernel"], "dram_bytes_per_launch": dram, "source": os.path.basename(rep), "tag": tag},
          open("profiles/coarse_kernel_summary.json", "w"), indent=1)

# source page: opcode mix + top stall sites of the dominant kernel
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
blocks, cur = [], None
for r in srows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": [], "hdr": None}; blocks.append(cur); continue
    if cur is None: continue
    if r and r[0] == "Address": cur["hdr"] = r; continue
    if cur["hdr"] and len(r) == len(cur["hdr"]): cur["rows"].append(r)
b = blocks[0]; h = b["hdr"]
si, ii, sc = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
stall_cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
tot_i = sum(int(r[ii]) for r in b["rows"]); tot_s = sum(int(r[si]) for r in b["rows"])
mix = collections.Counter()
for r in b["rows"]:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[sc])
    mix[(m.group(2).split(".")[0] if m else "?")] += int(r[ii])

with open(f"profiles/{tag}_coarse_kernel.md", "w") as f:
    f.write(f"# ncu --set full, dominant kernel ({tag})\n\nsource capture: `{os.path.basename(rep)}` "
            "(`ncu --set full --clock-control none --import-source on -k regex:coarse_kernel`)\n\n")
    for k in kernels:
        f.write(f"## {k['kernel']}\n\n| metric | value |\n|---|---|\n")
        for m in WANT:
            if m in k: f.write(f"| {m} | {k[m]} |\n")
        f.write("\n")
    f.write(f"## instruction mix of {b['name']}\n\ntotal warp instructions {tot_i}, samples {tot_s}\n\n| opcode | executed | share |\n|---|---|---|\n")
    for op, c in mix.most_common(20): f.write(f"| {op} | {c} | {100 * c / tot_i:.1f}% |\n")
    f.write("\n## top stall sites (warp samples)\n\n| samples | share | executed | SASS | top stall reasons |\n|---|---|---|---|---|\n")
    for idx, r in sorted(enumerate(b["rows"]), key=lambda t: -int(t[1][si]))[:25]:
        st = {h[j][6:]: int(r[j]) for j in stall_cols if int(r[j]) > 0}
        st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
        f.write(f"| {r[si]} | {100 * int(r[si]) / tot_s:.1f}% | {r[ii]} | `{r[sc].strip()[:70]}` | {st} |\n")

# launch list -> per-kernel totals of one step
seq, hd = [], None
for r in csv.reader(open(launches)):
    if "Kernel Name" in r: hd = r; ki, vi = hd.index("Kernel Name"), hd.index("Metric Value"); continue
    if hd and len(r) > vi and r[0].isdigit(): seq.append((r[ki], float(r[vi].replace(",", ""))))
names = [n for n, _ in seq]
# one step = from one codebook_prep to the next
starts = [i for i, n in enumerate(names) if "codebook_prep" in n]
step = seq[starts[-2]:starts[-1]] if len(starts) >= 2 else seq
tot = sum(v for _, v in step)
with open(f"profiles/{tag}_launches_one_step.md", "w") as f:
    f.write(f"# ncu launch list of one bench step ({tag})\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` "
            f"around `python bench.py --steps 2 --warmup 3 --no-cpu`; cold-cache, serialised: compare SHARES.\n\n"
            f"| kernel | us | share |\n|---|---|---|\n")
    for n, v in step: f.write(f"| `{n[:90]}` | {v / 1e3:.1f} | {100 * v / tot:.1f}% |\n")
    f.write(f"| **total** | {tot / 1e3:.1f} | |\n")
import shutil
shutil.copy(launches, f"profiles/{tag}_launches.csv")
print("dram bytes per launch of dominant kernel:", dram, "| one step (ncu) us:", tot / 1e3)
