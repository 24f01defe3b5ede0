"""Opcode histogram (instructions executed, stall samples) from `ncu -i rep --page source --csv --print-source sass`."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ops = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0]
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    src = r[ix["Source"]].strip()
    parts = src.split()
    op = parts[1] if parts and parts[0].startswith("@") and len(parts) > 1 else (parts[0] if parts else "?")
    op = op.split(".")[0]
    ex = int(r[ix["Instructions Executed"]] or 0)
    smp = int(r[ix["Warp Stall Sampling (All Samples)"]] or 0)
    ops[op][0] += ex; ops[op][1] += smp; ops[op][2] += 1
    tot[0] += ex; tot[1] += smp
print(f"static instrs {len(rows)-2}, executed {tot[0]}, samples {tot[1]}")
for op, (ex, smp, n) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:28]:
    print(f"{op:12s} exec {100*ex/tot[0]:5.1f}%  samples {100*smp/max(tot[1],1):5.1f}%  static {n}")
