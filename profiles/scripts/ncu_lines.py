import re,csv,sys,subprocess,os
rep,kname,mangled=sys.argv[1],sys.argv[2],sys.argv[3]
so='/root/repo/orb-ygz-slam_b200/libygz_b200.so'
os.system('rm -rf elf && mkdir elf && cd elf && cuobjdump -xelf all %s >/dev/null 2>&1'%so)
cub=[f for f in os.listdir('elf') if f.startswith(('orb_extractor','matcher','sparse'))]
li=None
for f in cub:
    out=subprocess.run(['nvdisasm','-g','-c','elf/'+f],capture_output=True,text=True).stdout
    if mangled in out: li=out; break
txt=li.split('\n')
start=[i for i,l in enumerate(txt) if '.text.'+mangled in l and '.section' in l][0]
cur=None; amap={}
for l in txt[start+1:]:
    if l.startswith('//---------------------') and amap: break
    m=re.search(r'//## File "(.*?)", line (\d+)',l)
    if m: cur=(os.path.basename(m.group(1)),int(m.group(2))); continue
    m=re.match(r'\s+/\*([0-9a-f]{4,5})\*/',l)
    if m and cur is not None: amap[int(m.group(1),16)]=cur
raw=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--kernel-name','regex:'+kname],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.split('\n')))
hi=[i for i,r in enumerate(rows) if 'Instructions Executed' in r][0]
hdr=rows[hi]; ie=hdr.index('Instructions Executed'); st=hdr.index('Warp Stall Sampling (All Samples)'); ad=hdr.index('Address')
f=lambda v: float(v) if v not in ('','-') else 0.0
base=None; per={}; pers={}
for r in rows[hi+1:]:
    if len(r)<=ie or not r[ad]: continue
    if 'Instructions Executed' in r: break   # second kernel instance
    a=int(r[ad],16) if r[ad].startswith('0x') else int(r[ad])
    if base is None: base=a
    line=amap.get(a-base,('?',-1))
    per[line]=per.get(line,0)+f(r[ie]); pers[line]=pers.get(line,0)+f(r[st])
tot=sum(per.values()); tots=sum(pers.values())
srcs={}
def srcline(fl):
    f,l=fl
    if f not in srcs:
        try: srcs[f]=open('/root/repo/orb-ygz-slam_b200/csrc/'+f).read().split('\n')
        except Exception: srcs[f]=None
    return srcs[f][l-1].strip()[:95] if srcs[f] and 0<l<=len(srcs[f]) else '?'
thr=float(sys.argv[4]) if len(sys.argv)>4 else 0.012
print('total warp inst',tot,'samples',tots)
for line in sorted(per):
    if per[line]>tot*thr or pers[line]>tots*thr:
        print('%-16s %5d %5.1f%% inst %5.1f%% stall | %s'%(line[0][:16],line[1],per[line]/tot*100,pers[line]/tots*100,srcline(line)))
