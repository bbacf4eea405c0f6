#!/usr/bin/env python
"""Per-SASS-region executed-instruction totals from an .ncu-rep source page: contiguous runs of
instructions with similar execution counts are merged, so loops/phases show up as regions."""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# a report holds one section per captured launch: pick the first whose kernel name contains argv[2] (default: first)
want = sys.argv[2] if len(sys.argv) > 2 else ""
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
sel = next((i for i in starts if want in rows[i][1]), starts[0])
end = next((i for i in starts if i > sel), len(rows))
print(rows[sel][1][:100])
rows = rows[sel:end]
hdr = rows[1]
ia, isrc, ie, it, isamp = (hdr.index(k) for k in ("Address", "Source", "Instructions Executed", "Thread Instructions Executed", "# Samples"))
ins = [(r[ia], r[isrc].strip(), float(r[ie] or 0), float(r[it] or 0), float(r[isamp] or 0)) for r in rows[2:] if len(r) > it]
tot = sum(x[2] for x in ins); tots = sum(x[4] for x in ins)
print(f"total warp instructions {tot:.0f}, samples {tots:.0f}")
regions = []
for i, x in enumerate(ins):
    if regions and 0.5 <= (x[2] + 1) / (regions[-1]["avg"] + 1) <= 2.0:
        r = regions[-1]; r["n"] += 1; r["sum"] += x[2]; r["thr"] += x[3]; r["samp"] += x[4]; r["avg"] = r["sum"] / r["n"]; r["end"] = i
    else:
        regions.append({"start": i, "end": i, "n": 1, "sum": x[2], "thr": x[3], "samp": x[4], "avg": x[2]})
for r in regions:
    if r["sum"] / tot < 0.004:
        continue
    ops = {}
    for x in ins[r["start"]:r["end"] + 1]:
        op = x[1].split()[0] if not x[1].startswith("@") else x[1].split()[1]
        op = op.split(".")[0]
        ops[op] = ops.get(op, 0) + 1
    top = ", ".join(f"{k}:{v}" for k, v in sorted(ops.items(), key=lambda t: -t[1])[:6])
    print(f"  instr #{r['start']:5d}-{r['end']:5d} ({r['n']:4d} instrs)  exec/instr {r['avg']:11.0f}  share {100*r['sum']/tot:5.1f}%  "
          f"lanes {r['thr']/max(r['sum'],1):4.1f}  samples {100*r['samp']/max(tots,1):5.1f}%  [{top}]")
