#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, total and first durations (us)."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if 'Kernel Name' in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            v = float(d['Metric Value'].replace(',', ''))
        except ValueError:
            continue
        u = d['Metric Unit']
        if u == 'ns':
            v /= 1000
        elif u == 'ms':
            v *= 1000
        elif u == 's':
            v *= 1e6
        k = d['Kernel Name'][:78]
        a = agg.setdefault(k, [0, 0.0, []])
        a[0] += 1
        a[1] += v
        a[2].append(round(v, 1))
tot = sum(a[1] for a in agg.values())
print("# %d launches, %.1f us total (cold-cache, serialised: compare shares)" % (sum(a[0] for a in agg.values()), tot))
for k, (n, t, l) in sorted(agg.items(), key=lambda x: -x[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    print("%-78s n=%3d total_us=%9.1f share=%5.1f%%  %s" % (k, n, t, 100 * t / tot, l[:6]))
