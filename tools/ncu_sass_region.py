#!/usr/bin/env python
"""Print the SASS around the TMA / tcgen05 instructions of one kernel of an `ncu --page source --csv
--print-source sass` dump, with executed counts and the top stall reasons per line.
usage: ncu_sass_region.py src.csv kernel_idx [before] [after]"""
import csv, sys
path, ki = sys.argv[1], int(sys.argv[2])
before = int(sys.argv[3]) if len(sys.argv) > 3 else 60
after = int(sys.argv[4]) if len(sys.argv) > 4 else 30
kernels = []; cur = None
for row in csv.reader(open(path)):
    if not row: continue
    if row[0] == 'Kernel Name': cur = {'name': row[1], 'hdr': None, 'rows': []}; kernels.append(cur)
    elif row[0] == 'Address': cur['hdr'] = row
    elif cur is not None and cur['hdr'] is not None: cur['rows'].append(row)
k = kernels[ki]; h = k['hdr']
iS, iN, iE = h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
stall_cols = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
rows = k['rows']
idx = [i for i, r in enumerate(rows) if any(t in r[iS] for t in ('UTCHMMA', 'UTMALDG', 'UTCBAR', 'UTCQMMA'))]
print(k['name'][:100], len(rows), 'lines')
lo = max(0, min(idx) - before); hi = min(len(rows), max(idx) + after)
for r in rows[lo:hi]:
    st = sorted(((int(float(r[i] or 0)), h[i][6:]) for i in stall_cols), reverse=True)[:2]
    print(r[0][-5:], f"{int(float(r[iE] or 0)):>9}", f"{int(float(r[iN] or 0)):>6}", r[iS].strip()[:90], ' '.join(f'{c}={n}' for n, c in st if n))
