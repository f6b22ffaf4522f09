"""Per source line: executed warp instructions and warp-stall samples of one kernel (diagnostic, no GPU needed).

Joins the SASS source page of an ncu report with the line table of the same build:

    ncu -i gpurun_out/x.ncu-rep --page source --csv > /tmp/src.csv
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -cubin -o /tmp/k.cubin lepton_b200/csrc/lep_capi.cu
    nvdisasm -g -c /tmp/k.cubin > /tmp/sass.txt
    python tests/tools_ncu_by_line.py /tmp/src.csv /tmp/sass.txt lep_decode_g2_kernelILi4E 60

(the cubin must be the build that was profiled: the two instruction lists are matched by position).  Columns: share of
all stall samples, share of executed instructions, average active lanes, samples by stall reason (long scoreboard = global
memory, wait = fixed-latency dependency, branch resolving, short scoreboard = shared memory)."""
import csv, re, sys, collections
ncu_csv, sass_txt, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
# nvdisasm: list of (file,line) per instruction in order
lines=open(sass_txt).read().splitlines()
start=None
for i,l in enumerate(lines):
    if l.startswith('//--------------------- .text.') and kern in l: start=i; break
assert start is not None
seq=[]; cur=('?',0)
for l in lines[start+1:]:
    if l.startswith('//--------------------- '): break
    m=re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)',l)
    if m:
        cur=(m.group(1).split('/')[-1],int(m.group(2))); 
        # inlined at?
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S',l):
        seq.append(cur)
rows=list(csv.reader(open(ncu_csv)))
hi=[i for i,r in enumerate(rows[:5]) if 'Source' in r and 'Address' in r][0]
hdr=rows[hi]
col={h:i for i,h in enumerate(hdr)}
body=[r for r in rows[hi+1:] if len(r)==len(hdr)]
print('ncu instr rows',len(body),'nvdisasm instrs',len(seq))
n=min(len(body),len(seq))
agg=collections.defaultdict(lambda:[0,0,0,0,0,0,0])
tot_i=tot_s=0
for k in range(n):
    r=body[k]; key=seq[k]
    ie=int(r[col['Instructions Executed']]); sm=int(r[col['# Samples']])
    a=agg[key]; a[0]+=ie; a[1]+=sm; a[2]+=int(r[col['stall_long_sb']]); a[3]+=int(r[col['stall_wait']]); a[4]+=int(r[col['stall_branch_resolving']]); a[5]+=int(r[col['stall_short_sb']]); a[6]+=int(r[col['Thread Instructions Executed']])
    tot_i+=ie; tot_s+=sm
print('total instr %d samples %d'%(tot_i,tot_s))
src={}
def srcline(f,n):
    import os
    for d in ('/root/repo/lepton_b200/csrc/',):
        p=d+f
        if os.path.exists(p):
            if p not in src: src[p]=open(p).read().splitlines()
            return src[p][n-1].strip()[:90] if n-1<len(src[p]) else ''
    return ''
for key,a in sorted(agg.items(),key=lambda kv:-kv[1][1])[:top]:
    print('%5.1f%% smp %5.1f%% inst (lanes %4.1f) long %6d wait %6d br %6d short %6d | %s:%d  %s'%(100*a[1]/tot_s,100*a[0]/tot_i,a[6]/max(a[0],1),a[2],a[3],a[4],a[5],key[0],key[1],srcline(*key)))
