"""Summarise an ncu source-page CSV: top stalled SASS instructions and the stall mix."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = rows[1]; ix = {k: i for i, k in enumerate(hdr)}; data = rows[2:]
tot = sum(int(r[ix['# Samples']] or 0) for r in data)
stall_cols = [k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]
print('total samples', tot, 'instrs', len(data), 'warp-inst', sum(int(r[ix['Instructions Executed']] or 0) for r in data))
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']] or 0))[:top_n]:
    n = int(r[ix['# Samples']])
    st = sorted(((int(r[ix[k]] or 0), k[6:]) for k in stall_cols), reverse=True)[:2]
    print(f"{n:6d} {100*n/tot:5.1f}% {r[ix['Address']][-5:]} {r[ix['Source']][:60]:60s} {st}")
agg = {k: sum(int(r[ix[k]] or 0) for r in data) for k in stall_cols}
print({k[6:]: round(100 * v / tot, 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]})
# opcode histogram by executed instructions
ops = {}
for r in data:
    src = r[ix['Source']].split()
    if not src: continue
    op = src[1] if src[0].startswith('@') and len(src) > 1 else src[0]
    op = op.split('.')[0]
    ops[op] = ops.get(op, 0) + int(r[ix['Instructions Executed']] or 0)
t = sum(ops.values())
print('executed opcode mix:', ', '.join(f"{k}:{100*v/t:.1f}%" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:28]))
