#!/usr/bin/env python
"""Turn an `ncu --set full --import-source on` report into the two text artefacts kept under profiles/:
   <out>_raw.csv      the raw metric page of the first profiled launch
   <out>_stalls.txt   per-phase and per-instruction warp-stall samples from the source page (SASS)
    python tools/ncu_export.py gpurun_out/prof_attn.ncu-rep profiles/r01_ncu_attn_final [--top 30]"""
import argparse, csv, io, subprocess

ap = argparse.ArgumentParser()
ap.add_argument("report"); ap.add_argument("out"); ap.add_argument("--top", type=int, default=30)
ap.add_argument("--kernel", default="", help="regex on the kernel name (reports holding several kernels)")
a = ap.parse_args()
sel = ["--kernel-name", "regex:" + a.kernel] if a.kernel else []
raw = subprocess.run(["ncu", "-i", a.report, *sel, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
with open(a.out + "_raw.csv", "w") as f:
    w = csv.writer(f)
    for k, unit, v in zip(rows[0], rows[1], rows[2]):
        w.writerow([k, v, unit])
src = subprocess.run(["ncu", "-i", a.report, *sel, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(src)))
name, hdr = rows[0][1], rows[1]
data = [r for r in rows[2:] if len(r) >= len(hdr) and r[0].startswith("0x")]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
S = lambda r: int(r[ix["# Samples"]])
tot = sum(S(r) for r in data)
lines = [f"{name}: {tot} warp-stall samples over {len(data)} SASS instructions"]
agg = {c: sum(int(r[ix[c]]) for r in data) for c in stalls}
lines.append("by reason: " + ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.005 * tot))
lines.append(f"top {a.top} instructions (samples, executed, SASS, reasons)")
for r in sorted(data, key=lambda r: -S(r))[:a.top]:
    why = {c[6:]: int(r[ix[c]]) for c in stalls if int(r[ix[c]]) > 0.1 * S(r)}
    lines.append(f"{S(r):6d} {int(r[ix['Instructions Executed']]):10d}  {r[ix['Source']].strip()[:70]:70s} {why}")
lines.append("samples per 64-instruction window (offset, samples, markers)")
for lo in range(0, len(data), 64):
    hi = min(len(data), lo + 64)
    s = sum(S(r) for r in data[lo:hi])
    marks = sorted({data[i][ix["Source"]].split()[0] for i in range(lo, hi)
                    if any(t in data[i][ix["Source"]] for t in ("LDTM", "STTM", "EXIT", "BAR.SYNC", "UTMALDG", "UTCHMMA", "UTCBAR", "STG.E", "TRYWAIT", "MUFU"))})
    if s > 0.004 * tot:
        lines.append(f"{lo:6d} {s:7d}  {' '.join(marks)}")
open(a.out + "_stalls.txt", "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:8]))
