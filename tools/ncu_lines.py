"""Per source line (outermost frame in a given file) instruction and stall-sample shares."""
import csv, sys, re, subprocess, collections, io
rep, gi, ksub, fname = sys.argv[1:5]
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); hdr = rows[1]; data = rows[2:]
iE = hdr.index('Instructions Executed'); iS = hdr.index('# Samples')
lines = open(gi).read().split('\n')
start = next(i for i, l in enumerate(lines) if l.startswith('.text.') and ksub in l)
seq = []; chain = []; fresh = True
for l in lines[start + 1:]:
    if l.startswith('.text.') or l.startswith('//--------------------- .'): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if fresh: chain = []; fresh = False
        chain.append((m.group(1).split('/')[-1], int(m.group(2)))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: seq.append((m.group(2), tuple(chain))); fresh = True
assert len(seq) == len(data), (len(seq), len(data))
cnt = collections.Counter(); smp = collections.Counter(); tot = ts = 0
for (t, ch), r in zip(seq, data):
    n = int(r[iE] or 0); s = int(r[iS] or 0); tot += n; ts += s
    k = None
    for f, ln in reversed(ch):
        if f == fname: k = ln; break
    if k is not None: cnt[k] += n; smp[k] += s
srcl = open('/root/repo/mole_b200/csrc/' + fname).read().split('\n')
for k in sorted(cnt):
    if smp[k] * 200 > ts or cnt[k] * 200 > tot:
        print("%4d  instr %5.2f%%  samples %5.2f%%  | %s" % (k, 100 * cnt[k] / tot, 100 * smp[k] / ts, srcl[k - 1].strip()[:90]))
