import re,subprocess,sys
log=open('/root/repo/mole_b200/csrc/_obj/ptxas_mole_api.log').read()
pat=sys.argv[1] if len(sys.argv)>1 else 'sj_'
for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\nptxas info\s+: Function properties for \S+\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers(.*)", log):
    name=m.group(1)
    if pat not in name: continue
    dn=subprocess.run(['c++filt',name],capture_output=True,text=True).stdout.strip()[:60]
    print(dn, 'regs',m.group(5),'stack',m.group(2),'spill st/ld',m.group(3),m.group(4))
