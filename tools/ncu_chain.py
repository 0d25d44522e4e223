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
def classify(ch):
    files = [f for f, _ in ch]
    if 'mole_rng.cuh' in files: return 'rng'
    # outermost frame within mole_sj_move.cuh / mole_sj.cuh
    for f, ln in reversed(ch):
        if f == 'mole_sj_move.cuh':
            return 'move:%03d' % (ln // 10 * 10)
    FN=[(0,'hdr'),(91,'gsum'),(104,'pair'),(122,'phi'),(132,'gradlnD'),(140,'radial'),(150,'invert'),(198,'refresh_slot'),(222,'refresh'),(232,'init'),(271,'swap'),(287,'movev2'),(457,'measure'),(539,'const/setup/load/store'),(586,'sweep_moves'),(611,'evalk'),(649,'sweepk')]
    outer=None
    for f, ln in reversed(ch):
        if f == 'mole_sj.cuh':
            if ln >= 649:    # kernel frame: look for the callee frame
                continue
            name=[n for a,n in FN if a<=ln][-1]
            if name in ('measure','refresh','refresh_slot','init','invert','swap','sweep_moves'):
                return 'sj:'+name+(':%03d'%(ln//10*10) if name=='measure' else '')
            outer = outer or name
    if outer: return 'sj:'+outer
    for f, ln in reversed(ch):
        if f == 'mole_sj.cuh': return 'sjk:%04d' % (ln // 10 * 10)
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
        print("%-12s instr %5.1f%%  samples %5.1f%%  (math.cuh share of this bucket %4.1f%%)" % (k, 100 * cnt[k] / tot, 100 * smp[k] / ts, 100 * mathc[k] / max(cnt[k], 1)))
