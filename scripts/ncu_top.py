"""Summarise an ncu report: headline metrics + the SASS instructions with the most stall samples.
usage: python scripts/ncu_top.py gpurun_out/prof_xxx.ncu-rep [N]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
body = [r for r in rows[2:] if len(r) >= len(hdr) - 2]
ci = hdr.index("Warp Stall Sampling (All Samples)")
ie = hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(float(r[ci] or 0) for r in body)
print(f"kernel: {rows[0][1]}\ntotal samples {tot:.0f}, SASS instructions {len(body)}")
agg = {}
for r in body:
    for i in stall_cols:
        agg[hdr[i]] = agg.get(hdr[i], 0) + float(r[i] or 0)
print("stall mix:", ", ".join(f"{k[6:]}={v/tot*100:.0f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]))
for r in sorted(body, key=lambda r: -float(r[ci] or 0))[:N]:
    st = sorted(((float(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
    print(f"{float(r[ci])/tot*100:5.1f}%  exec={int(float(r[ie] or 0)):>9d}  {r[1].strip()[:70]:70s} {st[0][1]}:{st[0][0]:.0f} {st[1][1]}:{st[1][0]:.0f}")
