#!/usr/bin/env python
"""Attribute ncu per-SASS-instruction counters to CUDA source lines (ncu's CSV source page has no CUDA view).
   usage: python tools/sass_lines.py <ncu source csv> <nvdisasm -g -c listing> <mangled function substring>"""
import collections, csv, re, sys
src_csv, dis, func = sys.argv[1:4]
lines = open(dis).read().split('\n')
start = next(i for i, l in enumerate(lines) if l.startswith('\t.section') and func in l and '.text' in l)
cur, instr = None, []
for l in lines[start + 1:]:
    if l.startswith('\t.section') or l.startswith('//-----'):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)), (m.group(3).split('/')[-1], int(m.group(4))) if m.group(3) else None)
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
        instr.append(cur)
rows = list(csv.reader(open(src_csv)))
hdr, data = rows[1], rows[2:]
iE, iSm = hdr.index('Instructions Executed'), hdr.index('# Samples')
assert len(data) == len(instr), (len(data), len(instr))
tot = sum(int(r[iE]) for r in data); ts = sum(int(r[iSm]) for r in data)
agg, aggs = collections.Counter(), collections.Counter()
for k, key in enumerate(instr):
    f, ln, inl = key or ('?', 0, None)
    lab = f"{f}:{ln}" + (f" <- {inl[0]}:{inl[1]}" if inl else "")
    agg[lab] += int(data[k][iE]); aggs[lab] += int(data[k][iSm])
print(f"total warp-instructions {tot}, samples {ts}")
for lab, e in agg.most_common(int(sys.argv[4]) if len(sys.argv) > 4 else 50):
    print(f"{e/tot*100:5.2f}% instr {aggs[lab]/ts*100:5.2f}% samp  {lab}")
