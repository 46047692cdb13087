"""Instruction mix / hottest SASS of one kernel from `ncu -i X.ncu-rep --page source --csv --kernel-name regex:K` output.
   python tools/ncu_hot.py file.csv [mix|top|lines]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) >= len(hdr) - 2 and r[idx['Instructions Executed']].isdigit()]
ie = lambda r: int(r[idx['Instructions Executed']])
smp = lambda r: int(r[idx['# Samples']] or 0)
tot = sum(ie(r) for r in data); ts = sum(smp(r) for r in data)
mx = max(ie(r) for r in data)
mode = sys.argv[2] if len(sys.argv) > 2 else 'mix'
print('total warp instr', tot, 'samples', ts, 'static instrs', len(data), 'max exec', mx)
def opname(r):
    t = r[idx['Source']].split()
    op = t[1] if t[0].startswith('@') else t[0]
    return op.split('.')[0]
if mode == 'mix':
    ops = collections.Counter(); sm = collections.Counter()
    for r in data: ops[opname(r)] += ie(r); sm[opname(r)] += smp(r)
    for o, c in ops.most_common(28): print(f"{o:10s} {100*c/tot:5.1f}% instr   {100*sm[o]/max(ts,1):5.1f}% samples")
elif mode == 'top':
    for r in sorted(data, key=lambda r: -smp(r))[:40]: print(f"{100*smp(r)/ts:5.1f}% smp {100*ie(r)/tot:5.2f}% ins  {r[idx['Source']][:90]}")
