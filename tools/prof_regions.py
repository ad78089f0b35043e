#!/usr/bin/env python
"""Per-function stall breakdown of one ncu capture (source page joined with nvdisasm -gi line info of the same cubin).
usage: prof_regions.py report.ncu-rep libscpp_b200.so '<kernel substring>' [ipm.cuh]"""
import csv, re, sys, collections, subprocess, os, tempfile
rep, so, ksub = sys.argv[1:4]
srcfile = sys.argv[4] if len(sys.argv) > 4 else os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'scpp_b200', 'csrc', 'ipm.cuh')
tmp = tempfile.mkdtemp()
# the library is linked from several objects whose embedded cubins share one name: extract each object on its own
objs = sorted(os.path.join(os.path.dirname(os.path.abspath(so)), '_obj', f) for f in os.listdir(os.path.join(os.path.dirname(os.path.abspath(so)), '_obj')) if f.endswith('.o'))
for i, o in enumerate(objs):
    subprocess.check_call(f"mkdir -p {tmp}/o{i} && cd {tmp}/o{i} && cuobjdump -xelf all {o} > /dev/null && for f in *.cubin; do nvdisasm -gi -c $f; done >> {tmp}/dis.txt", shell=True)
subprocess.check_call(f"ncu -i {rep} --page source --print-source sass --csv > {tmp}/sass.csv 2>/dev/null", shell=True)
subprocess.check_call(f"ncu -i {rep} --page raw --csv > {tmp}/raw.csv 2>/dev/null", shell=True)
# raw metrics
rows = list(csv.reader(open(f"{tmp}/raw.csv"))); h, u, v = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'smsp__inst_executed.sum', 'sm__warps_active.avg.per_cycle_active', 'smsp__issue_active.avg.pct', 'sm__inst_executed_pipe_fp64.avg.pct',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__shared_mem_per_block_dynamic', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__inst_executed_op_local', 'l1tex__t_bytes_pipe_lsu_mem_local']
for i, n in enumerate(h):
    if any(n == w or (n.startswith(w) and n.count('.') <= w.count('.') + 0) for w in want): print(f"{n} [{u[i]}] = {v[i]}")
kern = None; amap = {}; frames = []; fresh = True; on = False
for ln in open(f"{tmp}/dis.txt"):
    if ln.startswith('.text.'):
        on = ksub in ln; frames = []; continue
    if not on: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        if fresh: frames = []; fresh = False
        frames.append((m.group(1).split('/')[-1], int(m.group(2)))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', ln)
    if m: fresh = True; amap[int(m.group(1), 16)] = list(frames)
src = open(srcfile).read().split('\n'); fname = os.path.basename(srcfile)
funcs = []
for i, l in enumerate(src, 1):
    m = re.match(r'\s+(?:template <[^>]*>\s*)?SCPP_HD(?:_PASS|_CHOL)?\s+(?:static\s+)?(?:constexpr\s+)?[\w:<>\*& ]+?\s+\*?(\w+)\(', l)
    if m and not l.strip().startswith('//'): funcs.append((i, m.group(1)))
def fn_of(n):
    name = '?'
    for i, nm in funcs:
        if i <= n: name = nm
        else: break
    return name
TOP = {'pass_update', 'pass_residuals', 'pass_rhs', 'pass_recover', 'phase_factor', 'chain_forward', 'chain_backward', 'chol_inv', 'build_model_terms', 'solve', 'phase_solve', 'tables_init'}
def region(fr):
    # outermost frame inside ipm.cuh that belongs to a top-level phase
    for f, n in fr:
        if f == fname and fn_of(n) in TOP and fn_of(n) not in ('solve', 'phase_solve'): return fn_of(n)
    for f, n in fr:
        if f == fname: return fn_of(n)
    return fr[0][0] if fr else '?'
rows = list(csv.reader(open(f"{tmp}/sass.csv"))); hdr = rows[1]
ia, ism, iex = hdr.index('Address'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
cols = {n: hdr.index(n) for n in ['stall_long_sb', 'stall_wait', 'stall_short_sb', 'stall_selected', 'stall_branch_resolving', 'stall_no_inst', 'stall_lg', 'stall_mio']}
isrc = hdr.index('Source')
agg = collections.defaultdict(collections.Counter); base = None; line_s = collections.Counter()
for r in rows[2:]:
    if len(r) <= ism: continue
    a = int(r[ia], 16)
    if base is None: base = a
    fr = amap.get(a - base, [])
    reg = region(fr)
    for n, c in cols.items(): agg[reg][n] += int(r[c] or 0)
    agg[reg]['all'] += int(r[ism] or 0); agg[reg]['ex'] += int(r[iex] or 0); agg[reg]['static'] += 1
    op = r[isrc].split()[1] if r[isrc].strip().startswith('@') else r[isrc].split()[0]
    if op.startswith('LDL') or op.startswith('STL'): agg[reg]['local'] += int(r[iex] or 0)
    inner = next(((f, n) for f, n in fr if f == fname), None)
    if inner: line_s[inner[1]] += int(r[ism] or 0)
T = sum(v['all'] for v in agg.values()); E = sum(v['ex'] for v in agg.values())
print(f"total samples {T}, warp instructions {E}")
print(f"{'region':20s} smp%  ex%  long wait short sel  br noinst lg mio  local-ex%  code KB (static SASS x 16 B)")
for k, v in sorted(agg.items(), key=lambda t: -t[1]['all'])[:16]:
    a = v['all'] or 1
    print(f"{k:20s} {100*a/T:5.1f} {100*v['ex']/E:5.1f} " + ' '.join(f"{100*v[c]/a:4.0f}" for c in cols) + f"  {100*v['local']/max(1,v['ex']):5.1f}  {v['static']*16/1024:7.1f}")
print("top lines:")
for n, c in line_s.most_common(14): print(f"  {n:5d} {100*c/T:5.1f}  {src[n-1].strip()[:110]}")
