#!/usr/bin/env python
"""usage: ncu_analyze.py report.ncu-rep mangled_kernel_substring [topN]"""
import csv, sys, re, subprocess, collections, io, os
rep, ksub = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw))); hdr=rows[0]; d=dict(zip(hdr,rows[2])); units=dict(zip(hdr,rows[1]))
keys=['gpu__time_duration.sum','smsp__inst_executed.sum','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','dram__bytes_read.sum','dram__bytes_write.sum','l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_local_op_st.sum','smsp__inst_executed_op_shared_ld.sum','smsp__inst_executed_op_shared_st.sum']
print(d['Kernel Name'])
for k in keys:
    if k in d: print('  %-75s %s %s' % (k, d[k], units.get(k,'')))
src = subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(src))); hdr=rows[1]; data=rows[2:]
iS=hdr.index('Source'); iE=hdr.index('Instructions Executed'); iSamp=hdr.index('# Samples')
stall=['stall_wait','stall_no_inst','stall_short_sb','stall_branch_resolving','stall_not_selected','stall_math','stall_selected','stall_mio','stall_dispatch','stall_long_sb','stall_barrier','stall_lg','stall_membar']
ist=[hdr.index(k) for k in stall]
byop=collections.Counter(); tot=0
def opof(t):
    p=t.split(); o=p[1] if p[0].startswith('@') else p[0]; return o.split('.')[0]
ex=[]
for r in data:
    try: n=int(r[iE])
    except: n=0
    ex.append(n); byop[opof(r[iS])]+=n; tot+=n
print("static instrs %d  executed %d" % (len(data), tot))
print("opcode mix: " + "  ".join("%s %.1f" % (o,100*n/tot) for o,n in byop.most_common(22)))
s=[sum(int(r[i] or 0) for r in data) for i in ist]; T=sum(s)
print("stalls: " + "  ".join("%s %.1f" % (n[6:],100*v/T) for n,v in sorted(zip(stall,s),key=lambda t:-t[1])))
# line attribution
so=[f for f in os.popen("ls /root/repo/mole_b200/libmole_b200.so").read().split()]
os.system("rm -rf /tmp/xelf2 && mkdir /tmp/xelf2 && cd /tmp/xelf2 && cuobjdump -xelf all /root/repo/mole_b200/libmole_b200.so >/dev/null 2>&1 && nvdisasm -g -c mole_api.sm_100a.cubin > api.sass 2>/dev/null")
lines=open('/tmp/xelf2/api.sass').read().split('\n')
start=None
for i,l in enumerate(lines):
    if l.startswith('.text.') and ksub in l: start=i; break
seq=[]; cur=None
for l in lines[start+1:]:
    if l.startswith('//--------------------- .text') or (l.startswith('.text.') ): break
    m=re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur=(m.group(1).split('/')[-1], int(m.group(2))); continue
    m=re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: seq.append((m.group(2),cur))
print("sass instrs in cubin fn:", len(seq), "(ncu: %d)" % len(data))
if len(seq)==len(data):
    byline=collections.Counter()
    for (txt,cur),n in zip(seq,ex): byline[cur or ('?',0)]+=n
    srcs={}
    for (f,l),n in byline.most_common(topn):
        if f not in srcs:
            p='/root/repo/mole_b200/csrc/'+f
            srcs[f]=open(p).read().split('\n') if os.path.exists(p) else None
        t=srcs[f][l-1].strip()[:95] if srcs[f] and l-1<len(srcs[f]) else ''
        print("%5.2f%% %s:%d  %s" % (100*n/tot,f,l,t))

# ---- stall samples by source line
if len(seq)==len(data):
    iA=hdr.index('Warp Stall Sampling (All Samples)') if 'Warp Stall Sampling (All Samples)' in hdr else iSamp
    bl=collections.Counter(); bls={}
    for (txt,cur),r in zip(seq,data):
        k=cur or ('?',0)
        n=int(r[iSamp] or 0); bl[k]+=n
        st=bls.setdefault(k,[0]*len(stall))
        for j,i in enumerate(ist): st[j]+=int(r[i] or 0)
    T=sum(bl.values())
    print("\nstall samples by source line (top 40):")
    for (f,l),n in bl.most_common(40):
        if f not in srcs:
            p='/root/repo/mole_b200/csrc/'+f
            srcs[f]=open(p).read().split('\n') if os.path.exists(p) else None
        t=srcs[f][l-1].strip()[:70] if srcs[f] and l-1<len(srcs[f]) else ''
        st=bls[(f,l)]; top=sorted(zip(stall,st),key=lambda t:-t[1])[:2]
        print("%5.2f%% %s:%d [%s] %s" % (100*n/T,f,l," ".join("%s=%d%%"%(a[6:],100*b/max(sum(st),1)) for a,b in top),t))
