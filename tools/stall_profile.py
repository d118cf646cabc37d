#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv` (warp-stall samples per SASS instruction): totals per stall reason,
samples along the hot path in blocks of instructions, and the most-sampled instructions.  usage: stall_profile.py src.csv [block]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
B = int(sys.argv[2]) if len(sys.argv) > 2 else 60
print(rows[0][1] if len(rows[0]) > 1 else rows[0])
hdr = rows[1]
data = [dict(zip(hdr, r)) for r in rows[2:]]
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(d['# Samples']) for d in data)
print('total samples', tot, 'SASS instructions', len(data))
agg = {s: sum(int(d[s]) for d in data) for s in stalls}
print(' '.join('%s:%.1f%%' % (k[6:], 100 * v / tot) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:12]))
mx = max(int(d['Instructions Executed']) for d in data)
hot = [(i, d) for i, d in enumerate(data) if int(d['Instructions Executed']) >= 0.2 * mx]
print('per-quad instructions (executed >= 20%% of max %d): %d, samples there: %d (%.1f%%)' % (mx, len(hot), sum(int(d['# Samples']) for i, d in hot), 100 * sum(int(d['# Samples']) for i, d in hot) / tot))
for b in range(0, len(hot), B):
    blk = hot[b:b + B]
    s = sum(int(d['# Samples']) for i, d in blk)
    st = collections.Counter(); ops = collections.Counter()
    for i, d in blk:
        for k in stalls: st[k[6:]] += int(d[k])
        op = [o for o in d['Source'].split() if not o.startswith('@')][0].split('.')[0]
        ops[op] += 1
    print('%5d-%5d samples %4d (%4.1f%%) | %-60s | %s' % (blk[0][0], blk[-1][0], s, 100 * s / tot, ' '.join('%s:%d' % (k, v) for k, v in st.most_common(4)), ' '.join('%s:%d' % (k, v) for k, v in ops.most_common(6))))
print('--- most sampled instructions')
for i, d in sorted(enumerate(data), key=lambda x: -int(x[1]['# Samples']))[:24]:
    st = sorted(((int(d[k]), k[6:]) for k in stalls), reverse=True)[:2]
    print('%5d exec %7s samples %4s %-40s %s' % (i, d['Instructions Executed'], d['# Samples'], st, d['Source'].strip()[:80]))
