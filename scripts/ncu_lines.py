"""Per-source-line roll-up of an ncu report: instructions executed and stall samples.

    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --launch-skip 0 --launch-count 1 > cs.csv
    python scripts/ncu_lines.py cs.csv [top_n]
"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path, newline="")))
cur_file, hdr = None, None
inst = defaultdict(float)
samp = defaultdict(float)
stall = defaultdict(lambda: defaultdict(float))
src = {}
smem_wave = defaultdict(float)
smem_ideal = defaultdict(float)
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0] == "":
        continue  # SASS rows under a source line: the line row already carries their sum
    d = dict(zip(hdr[2:], r[2:]))  # skip the duplicated "Source" header of the cuda column
    key = (cur_file, int(r[0]))
    src.setdefault(key, r[1])
    try:
        inst[key] += float(d["Instructions Executed"])
        samp[key] += float(d["# Samples"])
        smem_wave[key] += float(d["L1 Wavefronts Shared"])
        smem_ideal[key] += float(d["L1 Wavefronts Shared Ideal"])
        for k, v in d.items():
            if k.startswith("stall_") and not k.endswith("(Not Issued)") and float(v) > 0:
                stall[key][k] += float(v)
    except (KeyError, ValueError):
        continue
ti, ts = sum(inst.values()), sum(samp.values())
print(f"total warp-instructions {ti:.3e}   samples {ts:.0f}")
tot_stall = defaultdict(float)
for k in stall:
    for s, v in stall[k].items():
        tot_stall[s] += v
print("stall mix:", ", ".join(f"{s[6:]} {100*v/ts:.1f}%" for s, v in sorted(tot_stall.items(), key=lambda x: -x[1])[:10]))
print(f"\n--- top {top} lines by samples ---")
for key in sorted(samp, key=lambda k: -samp[k])[:top]:
    st = ", ".join(f"{s[6:]} {v:.0f}" for s, v in sorted(stall[key].items(), key=lambda x: -x[1])[:3])
    print(f"{key[0]}:{key[1]:4d}  samp {100*samp[key]/ts:5.1f}%  inst {100*inst[key]/ti:5.1f}%  smem {smem_wave[key]:.2e}/{smem_ideal[key]:.2e}  [{st}]  {src[key].strip()[:90]}")
print(f"\n--- top {top} lines by instructions ---")
for key in sorted(inst, key=lambda k: -inst[k])[:top]:
    print(f"{key[0]}:{key[1]:4d}  inst {100*inst[key]/ti:5.1f}%  samp {100*samp[key]/ts:5.1f}%  {src[key].strip()[:100]}")

# optional: roll-up by line ranges "name:lo-hi,name:lo-hi" (third argument), kernels.cu only
if len(sys.argv) > 3:
    print("\n--- regions ---")
    for spec in sys.argv[3].split(","):
        name, rng = spec.split(":")
        lo, hi = [int(x) for x in rng.split("-")]
        i = sum(v for (f, l), v in inst.items() if f == "kernels.cu" and lo <= l <= hi)
        s = sum(v for (f, l), v in samp.items() if f == "kernels.cu" and lo <= l <= hi)
        print(f"{name:14s} inst {100*i/ti:5.1f}%  samp {100*s/ts:5.1f}%")
    i = sum(v for (f, l), v in inst.items() if f != "kernels.cu")
    s = sum(v for (f, l), v in samp.items() if f != "kernels.cu")
    print(f"{'other files':14s} inst {100*i/ti:5.1f}%  samp {100*s/ts:5.1f}%")
