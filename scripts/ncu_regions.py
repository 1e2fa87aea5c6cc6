"""Summarise `ncu --page source --csv` output: per kernel, the instructions with the most stall samples and the
samples grouped by code region (runs of instructions with similar execution counts).
    ncu -i rep.ncu-rep --page source --csv > src.csv ; python scripts/ncu_regions.py src.csv [kernel-substring] [top]"""
import csv, math, sys
path = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ""; top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
rows = list(csv.reader(open(path)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
seen = set()
for si, s in enumerate(starts):
    name = rows[s][1]
    if want not in name or name in seen:
        continue
    seen.add(name)
    end = starts[si + 1] if si + 1 < len(starts) else len(rows)
    H = rows[s + 1]; data = [r for r in rows[s + 2:end] if len(r) == len(H)]
    isrc, isamp, iex, ithr = H.index("Source"), H.index("# Samples"), H.index("Instructions Executed"), H.index("Thread Instructions Executed")
    stall = [(i, h) for i, h in enumerate(H) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[isamp]) for r in data); ins = sum(int(r[iex]) for r in data); thr = sum(int(r[ithr]) for r in data)
    print(f"=== {name[:100]}\n    SASS instructions {len(data)}, samples {tot}, warp instructions executed {ins}, thread instructions {thr} "
          f"({thr / max(ins, 1):.1f} lanes/instr)")
    agg = {}
    for r in data:
        for c, h in stall:
            if r[c]:
                agg[h] = agg.get(h, 0) + int(r[c])
    print("    stall mix:", ", ".join(f"{h[6:]} {100 * v / max(tot, 1):.1f}%" for h, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
    print("    -- top instructions by samples: index, samples, executed, SASS, top stalls")
    for i in sorted(sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:top]):
        r = data[i]
        st = sorted([(int(r[c]), h[6:]) for c, h in stall if r[c] and int(r[c]) > 0], reverse=True)[:3]
        print(f"    {i:5d} {int(r[isamp]):7d} {int(r[iex]):10d}  {r[isrc][:64]:64s} {st}")
    print("    -- regions (runs of similar execution count): [first, last) instrs samples share avg-executed")
    prev, a = None, 0
    runs = []
    for i, r in enumerate(data):
        e = int(r[iex]); key = 0 if e == 0 else round(math.log10(e) * 3)
        if key != prev:
            if prev is not None: runs.append((a, i))
            a, prev = i, key
    runs.append((a, len(data)))
    for a, b in runs:
        sm = sum(int(data[i][isamp]) for i in range(a, b)); ex = sum(int(data[i][iex]) for i in range(a, b))
        if sm > tot * 0.004: print(f"    [{a:5d},{b:5d}) {b - a:4d} {sm:7d} {100 * sm / tot:5.1f}%  {ex // (b - a):10d}")
