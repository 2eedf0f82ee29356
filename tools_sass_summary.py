"""SASS mnemonic counts of the shipped library, per kernel -> profiles/<tag>_sass_summary.md
    python tools_sass_summary.py <tag>
(cuobjdump -sass medtok_b200/libmedtok_vq.so; runs without a GPU).  Diagnostic tool."""
import collections
import re
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "sass"
txt = subprocess.run(["cuobjdump", "-sass", "medtok_b200/libmedtok_vq.so"], capture_output=True, text=True).stdout
WHOLE = ["UTCHMMA", "UTCBAR", "LDTM", "UBLKCP", "UTMALDG", "LDGSTS", "SYNCS", "FMNMX3", "F2FP", "STS", "RED", "ATOM", "HMMA", "FFMA", "BAR."]
PER = ["UTCHMMA", "UTCBAR", "LDTM", "UBLKCP", "LDGSTS", "UTMALDG", "SYNCS", "FMNMX3", "F2FP"]
kern, name = collections.OrderedDict(), None
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        kern[name] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        op = m.group(1)
        kern[name]["_n"] += 1
        for w in WHOLE:
            if op.startswith(w):
                kern[name][w] += 1
tot = collections.Counter()
for c in kern.values():
    tot.update(c)
out = [f"# SASS summary of the shipped library ({tag})", "",
       "`cuobjdump -sass medtok_b200/libmedtok_vq.so` (sm_100a cubin), mnemonic counts per kernel; `UTCHMMA` = tcgen05.mma, `LDTM` = tcgen05.ld,",
       "`UTCBAR` = tcgen05.commit, `UBLKCP` = cp.async.bulk (1-D bulk copy into the codebook ring; no tensor map is used, hence no `UTMALDG`),",
       "`LDGSTS` = cp.async (the REFINE pass's row gather into its staging tiles), `SYNCS` = mbarrier operations, `FMNMX3` = 3-input max of the",
       "epilogue's reduction trees.", "",
       "| whole library | " + " | ".join(WHOLE) + " |", "|---" * (len(WHOLE) + 1) + "|",
       "| total | " + " | ".join(str(tot[w]) for w in WHOLE) + " |", "",
       "| kernel | SASS instructions | " + " | ".join(PER) + " |", "|---" * (len(PER) + 2) + "|"]
for n, c in sorted(kern.items(), key=lambda t: -t[1]["_n"]):
    out.append(f"| `{n}` | {c['_n']} | " + " | ".join(str(c[w]) for w in PER) + " |")
open(f"profiles/{tag}_sass_summary.md", "w").write("\n".join(out) + "\n")
print("\n".join(out[:14]))
print("\n".join(l for l in out if "coarse_kernel<4, 5" in l or "coarse_kernel<1, 6" in l))
