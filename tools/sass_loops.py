import re,collections,sys,subprocess
fun=sys.argv[1]
out=subprocess.run(['cuobjdump','-sass','-fun',fun,(sys.argv[4] if len(sys.argv)>4 else "/root/repo/xpore_b200/libxpore_b200.so")],capture_output=True,text=True).stdout
ins=[]
for l in out.split('\n'):
    m=re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);',l)
    if m: ins.append((int(m.group(1),16),m.group(2).strip()))
print('total instr',len(ins))
loops=[]
for a,t in ins:
    if 'BRA' in t and 'BRA.DIV' not in t:
        mm=re.search(r'0x([0-9a-f]+)',t)
        if mm:
            tgt=int(mm.group(1),16)
            if tgt<a: loops.append((tgt,a))
def opname(t):
    parts=t.split()
    op=parts[1] if parts[0].startswith('@') else parts[0]
    return op.split('.')[0]
for tgt,a in loops:
    body=[(x,t) for x,t in ins if tgt<=x<=a]
    nf=sum(1 for x,t in body if opname(t) in('DFMA','DADD','DMUL','DSETP'))
    if nf>=20 or (len(sys.argv)>2):
        c=collections.Counter(opname(t) for x,t in body)
        print(hex(tgt),hex(a),'len',len(body),'fp64',nf,c.most_common(14))
if len(sys.argv)>2:
    lo,hi=int(sys.argv[2],16),int(sys.argv[3],16)
    for x,t in ins:
        if lo<=x<=hi: print(hex(x),t)
