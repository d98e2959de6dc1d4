"""Summarise an `ncu --page source --csv --print-source=sass` export: executed-instruction mix and hot blocks."""
import csv, collections, sys
path, npx = sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else None
rows = list(csv.reader(open(path)))
out = []
kernel = None
for r in rows:
    if r and r[0] == 'Kernel Name': kernel = r[1]; hdr = None; continue
    if r and r[0] == 'Address': hdr = r; first = not out; continue
    if kernel and hdr and len(r) == len(hdr): out.append((kernel, hdr, r))
k0 = out[0][0]
data = [(h, r) for k, h, r in out if k == k0]
hdr = data[0][0]
ia, ie = hdr.index('Source'), hdr.index('Instructions Executed')
# keep only first launch (addresses repeat)
seen, rows1 = set(), []
for h, r in data:
    if r[0] in seen: break
    seen.add(r[0]); rows1.append(r)
tot = sum(int(r[ie]) for r in rows1)
print('kernel', k0[:60], 'warp-instr', tot, 'static', len(rows1), ('lane-instr/px %.2f' % (tot * 32 / npx)) if npx else '')
h = collections.Counter()
for r in rows1:
    parts = r[ia].split()
    op = parts[1] if parts[0].startswith('@') else parts[0]
    h[op.split('.')[0]] += int(r[ie])
for op, c in h.most_common(22):
    print(f'{op:10s} {c:12d} {100*c/tot:5.1f}%' + (f'  /px {c*32/npx:.2f}' if npx else ''))
blocks, cur = [], None
for idx, r in enumerate(rows1):
    e = int(r[ie])
    if cur is None or e != cur[2]: cur = [idx, idx, e]; blocks.append(cur)
    else: cur[1] = idx
for b in sorted(sorted(blocks, key=lambda b: -(b[1]-b[0]+1)*b[2])[:14]):
    n = b[1]-b[0]+1
    print(f'block {b[0]:5d}-{b[1]:5d} n={n:4d} exec/instr={b[2]:9d} share={100*n*b[2]/tot:5.1f}%  first: {rows1[b[0]][ia][:50]}')
