import re,sys,subprocess
so=sys.argv[1]
def score(fun):
    out=subprocess.run(['cuobjdump','-sass','-fun',fun,so],capture_output=True,text=True).stdout
    ins=[]
    for l in out.split('\n'):
        m=re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);',l)
        if m: ins.append((int(m.group(1),16),m.group(2).strip()))
    # find main loop: backward branch whose body has 60-70 fp64 ops
    best=None
    for a,t in ins:
        if 'BRA' in t and 'DIV' not in t:
            mm=re.search(r'0x([0-9a-f]+)',t)
            if mm:
                tgt=int(mm.group(1),16)
                if tgt<a:
                    body=[(x,u) for x,u in ins if tgt<=x<=a]
                    nf=sum(1 for x,u in body if re.match(r'(@\S+\s+)?D(FMA|ADD|MUL)',u))
                    if 56<=nf<=140 and (best is None or len(body)<len(best)): best=body
    if best is None: return None
    fp=[u for x,u in best if re.match(r'(@\S+\s+)?D(FMA|ADD|MUL)',u)]
    dep=0
    for p,q in zip(fp,fp[1:]):
        d=re.search(r'D\w+\s+(R\d+)',p).group(1)
        srcs=re.findall(r'R\d+',q.split(',',1)[1])
        if d in srcs: dep+=1
    nops=sum(1 for x,u in best if u.startswith('NOP'))
    return len(best),len(fp),dep,nops
for f in ['_ZN3xpk13xp_fit_kernelILi1ELi512EEEvNS_7FitArgsE','_ZN3xpk13xp_fit_kernelILi1ELi640EEEvNS_7FitArgsE']:
    print(f[26:33], 'loop_len, fp64, adjacent_dependent_pairs, nops =', score(f))
