"""Summarise `ncu --page source --csv` output: opcode mix, stall mix, hottest instructions.
usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:K > src.csv; python tools/ncu_src.py src.csv [ntop] [launch]"""
import csv, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 20
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0  # launch index inside the file
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
a = starts[which]; b = starts[which + 1] if which + 1 < len(starts) else len(rows)
rows = rows[a:b]
hdr = rows[1]; idx = {n: i for i, n in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) >= len(hdr) and r[0] != "Address"]
tot = sum(int(r[idx['# Samples']]) for r in data)
inst = sum(int(r[idx['Instructions Executed']]) for r in data)
print('kernel', rows[0][1][:80]); print('samples', tot, 'inst', inst, 'sass lines', len(data))
ci = Counter(); cs = Counter()
for r in data:
    op = r[idx['Source']].split()
    o = op[0] if not op[0].startswith('@') else op[1]
    o = o.split('.')[0]
    ci[o] += int(r[idx['Instructions Executed']]); cs[o] += int(r[idx['# Samples']])
for o, c in ci.most_common(14):
    print(f"{o:10s} inst {c / inst * 100:5.1f}%  samples {cs[o] / tot * 100:5.1f}%")
stalls = [k for k in idx if k.startswith('stall_') and 'Not Issued' not in k]
sc = Counter()
for r in data:
    for k in stalls:
        sc[k] += int(r[idx[k]])
print({k: round(v / tot * 100, 1) for k, v in sc.most_common(8)})
for r in sorted(data, key=lambda r: -int(r[idx['# Samples']]))[:ntop]:
    st = {k[6:]: int(r[idx[k]]) for k in stalls if int(r[idx[k]]) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(r[idx['# Samples']], r[idx['Instructions Executed']], r[idx['Source']][:60], st)
