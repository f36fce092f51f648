import re, csv, collections, subprocess, sys, os
rep = sys.argv[1]; func = sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv)>3 else 30
os.system('cd /tmp && rm -rf xelf && mkdir xelf && cd xelf && cuobjdump -xelf all /root/repo/gpslam_b200/libgpb.so > /dev/null 2>&1 && nvdisasm -g engine.sm_100a.cubin > /tmp/all_dis.txt 2>/dev/null')
os.system('ncu -i %s --page source --csv --kernel-name regex:%s > /tmp/src_x.csv 2>/dev/null' % (rep, sys.argv[4] if len(sys.argv)>4 else '.'))
lines = open('/tmp/all_dis.txt').read().split('\n')
start = next(i for i,l in enumerate(lines) if l.startswith('.text.'+func+':'))
cur=None; seq=[]
for l in lines[start+1:]:
    if l.startswith('//-----') : break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur=(m.group(1).split('/')[-1], int(m.group(2))); continue
    m2 = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m2: seq.append((cur, m2.group(2).strip()))
rows = list(csv.reader(open('/tmp/src_x.csv')))
hdr = rows[1]; si = hdr.index('Source'); wi = hdr.index('# Samples'); ii = hdr.index('Instructions Executed')
data = rows[2:]
print("sass", len(seq), "ncu rows", len(data))
n = min(len(seq), len(data))
agg = collections.defaultdict(lambda:[0,0])
for k in range(n):
    agg[seq[k][0]][0]+=int(data[k][wi]); agg[seq[k][0]][1]+=int(data[k][ii])
tot = sum(v[0] for v in agg.values()); toti=sum(v[1] for v in agg.values())
src = {}
for f in os.listdir('/root/repo/gpslam_b200/csrc'):
    src[f] = open('/root/repo/gpslam_b200/csrc/'+f).read().split('\n')
print("total samples", tot, "instr", toti)
for key,v in sorted(agg.items(), key=lambda x:-x[1][0])[:top]:
    text = src[key[0]][key[1]-1].strip()[:95] if key and key[0] in src else ''
    print("%6d (%4.1f%%) instr %10d (%4.1f%%) | %s:%s  %s" % (v[0],100*v[0]/tot,v[1],100*v[1]/toti,key[0] if key else None,key[1] if key else None,text))
