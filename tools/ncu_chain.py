"""Attribute executed instructions / stall samples of one kernel to source frames using nvdisasm -gi chains.
usage: ncu_chain.py report.ncu-rep nvdisasm_gi.txt kernel_mangled_substr"""
import csv, sys, re, subprocess, collections, io
rep, gi, ksub = sys.argv[1:4]
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); hdr = rows[1]; data = rows[2:]
iE = hdr.index('Instructions Executed'); iS = hdr.index('# Samples')
lines = open(gi).read().split('\n')
start = next(i for i, l in enumerate(lines) if l.startswith('.text.') and ksub in l)
seq = []; chain = []
fresh = True
for l in lines[start + 1:]:
    if l.startswith('.text.') or l.startswith('//--------------------- .'): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if fresh: chain = []; fresh = False
        chain.append((m.group(1).split('/')[-1], int(m.group(2))))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        seq.append((m.group(2), tuple(chain))); fresh = True
assert len(seq) == len(data), (len(seq), len(data))
import os
SJ = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'mole_b200', 'csrc', 'mole_sj.cuh')
# line -> enclosing function of mole_sj.cuh, derived from the source (MOLE_D / __global__ / static inline definitions)
FN = [(0, 'hdr')]
for i, l in enumerate(open(SJ), 1):
    m = re.match(r'^(?:template.*>\s*)?(?:MOLE_D|__global__|__device__ __forceinline__|static inline)\b.*?\b(\w+)\(', l)
    if m and not l.startswith(' '): FN.append((i, m.group(1)))
def fn_of(ln): return [n for a, n in FN if a <= ln][-1]
PHASE = ('sj_measure', 'sj_refresh', 'sj_refresh_slot', 'sj_init', 'sj_invert', 'sj_swap_slots', 'sj_fold_psi')
def classify(ch):
    files = [f for f, _ in ch]
    if 'mole_rng.cuh' in files: return 'rng'
    for f, ln in reversed(ch):                       # outermost frame within mole_sj_move.cuh
        if f == 'mole_sj_move.cuh': return 'move:%03d' % (ln // 10 * 10)
    names = [fn_of(ln) for f, ln in ch if f == 'mole_sj.cuh']
    for n in PHASE:
        if n in names: return n
    if 'sj_sweep_moves' in names: return 'sj_sweep_moves (loop, draws)'
    if names: return names[-1]
    return files[-1] if files else '?'
cnt = collections.Counter(); smp = collections.Counter(); tot = 0; ts = 0
mathc = collections.Counter()
for (t, ch), r in zip(seq, data):
    n = int(r[iE] or 0); s = int(r[iS] or 0)
    k = classify(ch); cnt[k] += n; smp[k] += s; tot += n; ts += s
    if ch and ch[0][0] == 'mole_math.cuh': mathc[k] += n
print("total warp-instr %d samples %d" % (tot, ts))
for k in sorted(cnt):
    if cnt[k] * 1000 > tot or smp[k] * 1000 > ts:
        print("%-30s instr %5.1f%%  samples %5.1f%%  (math.cuh share of this bucket %4.1f%%)" % (k, 100 * cnt[k] / tot, 100 * smp[k] / ts, 100 * mathc[k] / max(cnt[k], 1)))
