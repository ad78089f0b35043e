#!/usr/bin/env python
"""Join an `ncu --page source --print-source sass --csv` export with `nvdisasm -g -c` line info of the same cubin and
aggregate stall samples / executed instructions by source line and by named source region.
usage: prof_by_line.py sass.csv dis.txt '<mangled kernel name>' source_file [regions.json]"""
import csv, re, sys, collections
sass_csv, dis, kern, srcfile = sys.argv[1:5]
# 1) address -> (file, line) from nvdisasm
amap = {}
cur = None; on = False
for ln in open(dis):
    if ln.startswith('.text.'):
        on = ln.strip() == '.text.' + kern + ':'
        continue
    if not on: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', ln)
    if m: amap[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
ia, isrc, ismp, iex = hdr.index('Address'), hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
base = None
by_line = collections.Counter(); ex_line = collections.Counter(); ops = collections.Counter(); opsmp = collections.Counter()
tot_s = tot_e = 0
for r in rows[2:]:
    if len(r) <= iex: continue
    a = int(r[ia], 16)
    if base is None: base = a
    loc = amap.get(a - base, ('?', 0))
    smp, ex = int(r[ismp] or 0), int(r[iex] or 0)
    by_line[loc] += smp; ex_line[loc] += ex; tot_s += smp; tot_e += ex
    op = r[isrc].split()[0] if not r[isrc].strip().startswith('@') else r[isrc].split()[1]
    op = op.split('.')[0]
    ops[op] += ex; opsmp[op] += smp
print(f"total samples {tot_s}  warp-instructions executed {tot_e}")
print("== top opcodes (executed %, samples %)")
for op, e in ops.most_common(18): print(f"  {op:10s} {100*e/tot_e:5.1f}  {100*opsmp[op]/tot_s:5.1f}")
# 2) regions of the source file by function: find 'SCPP_HD ... name(' lines
src = open(srcfile).read().split('\n')
fname = srcfile.split('/')[-1]
funcs = []
for i, l in enumerate(src, 1):
    m = re.match(r'\s*(?:template <[^>]*>\s*)?SCPP_HD\s+(?:static\s+)?(?:constexpr\s+)?[\w:<>\*& ]+?\s+\*?(\w+)\(', l)
    if m and not l.strip().startswith('//'): funcs.append((i, m.group(1)))
def region(loc):
    f, n = loc
    if f != fname: return f
    name = '?'
    for i, nm in funcs:
        if i <= n: name = nm
        else: break
    return name
reg_s = collections.Counter(); reg_e = collections.Counter()
for loc, v in by_line.items(): reg_s[region(loc)] += v
for loc, v in ex_line.items(): reg_e[region(loc)] += v
print("== by function/region (samples %, executed %)")
for k, v in reg_s.most_common(30): print(f"  {k:22s} {100*v/tot_s:5.1f}  {100*reg_e[k]/tot_e:5.1f}")
print("== top source lines (samples %, executed %)")
for loc, v in by_line.most_common(45):
    f, n = loc
    text = src[n-1].strip()[:110] if f == fname and 0 < n <= len(src) else ''
    print(f"  {f}:{n:5d} {100*v/tot_s:5.1f} {100*ex_line[loc]/tot_e:5.1f}  {text}")
if len(sys.argv) > 5:
    lo, hi = map(int, sys.argv[5].split('-'))
    print(f"== lines {lo}-{hi} of {fname} by executed instructions (executed %, samples %)")
    sel = [(loc, e) for loc, e in ex_line.items() if loc[0] == fname and lo <= loc[1] <= hi]
    for loc, e in sorted(sel, key=lambda t: -t[1])[:40]:
        print(f"  {loc[1]:5d} {100*e/tot_e:5.1f} {100*by_line[loc]/tot_s:5.1f}  {src[loc[1]-1].strip()[:120]}")
