"""Lock-step efficiency model for lep_decode_lockstep.cu (diagnostic, CPU only; needs PIL for bench.synth_jpeg).

Counts, per block of 8 bench images x 4 segments = 32 lanes, the decisions of the three step phases (non-zero count + 7x7,
edge counts + edges, DC) straight from the coefficient planes, lines the lanes up block by block in decode order and
reports mean steps per block against the warp cost (max over the 32 lanes) -- the SIMT efficiency quoted in DESIGN.md
section 4 -- and the bound a scheme without per-block meeting points could reach."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from lepton_b200 import HostJpeg

def bitlen(a):
    a=np.asarray(a,dtype=np.int64); out=np.zeros(a.shape,dtype=np.int64)
    nzm=a>0; out[nzm]=np.floor(np.log2(a[nzm])).astype(np.int64)+1; return out
def coef_cost(v):
    a=np.abs(v.astype(np.int64)); l=bitlen(a)
    c=np.where(a==0,1,l+(l<11)+1+np.maximum(l-1,0))
    return c
def block_steps(p):
    # p: [n,64] aligned order. returns (A,B,C) step counts per block
    n=p.shape[0]
    c77=coef_cost(p[:,:49]); nz=(p[:,:49]!=0)
    # positions up to last nonzero
    idx=np.arange(49)[None,:]
    last=np.where(nz.any(1), 48-np.argmax(nz[:,::-1],axis=1), -1)
    A=6+np.where(idx<=last[:,None],c77,0).sum(1)
    B=np.zeros(n,dtype=np.int64)
    for lo in (50,57):
        e=p[:,lo:lo+7]; ce=coef_cost(e); nze=e!=0
        laste=np.where(nze.any(1), 6-np.argmax(nze[:,::-1],axis=1), -1)
        B+=3+np.where(np.arange(7)[None,:]<=laste[:,None],ce,0).sum(1)
    dc=p[:,49].astype(np.int64); d=np.diff(dc,prepend=dc[:1])
    C=coef_cost(d)
    return A,B,C

lanes=[]
for seed in range(8):
    hj=HostJpeg(bench.synth_jpeg(seed)); img=hj.coef_image()
    starts=list(img.luma_y_start)+[img.bcv[0]]
    steps=[block_steps(np.asarray(pl).reshape(-1,64)) for pl in img.planes]
    v0=img.bcv[0]//img.mcuv
    for s in range(len(starts)-1):
        seqA=[];seqB=[];seqC=[]
        for mrow in range(starts[s]//v0, starts[s+1]//v0):
            for c in (2,1,0):
                mult=img.bcv[c]//img.mcuv; w=img.bch[c]
                for r in range(mrow*mult,(mrow+1)*mult):
                    sl=slice(r*w,(r+1)*w)
                    seqA.append(steps[c][0][sl]);seqB.append(steps[c][1][sl]);seqC.append(steps[c][2][sl])
        lanes.append([np.concatenate(x) for x in (seqA,seqB,seqC)])
print('lanes',len(lanes),'blocks per lane',[len(l[0]) for l in lanes][:8])
n=min(len(l[0]) for l in lanes); N=max(len(l[0]) for l in lanes)
tot_mean=0; tot_max=0
for ph,name in enumerate('ABC'):
    M=np.zeros((len(lanes),N),dtype=np.int64)
    for i,l in enumerate(lanes): M[i,:len(l[ph])]=l[ph]
    mean=M.sum()/len(lanes)/N; mx=M.max(0).mean()
    tot_mean+=mean; tot_max+=mx
    print('phase',name,'mean steps/block %.1f'%mean,'warp steps/block-round (max over 32 lanes) %.1f'%mx,'efficiency %.2f'%(mean/mx))
print('total mean %.1f max %.1f eff %.2f'%(tot_mean,tot_max,tot_mean/tot_max))
# alternative: no per-block sync (ideal free-running) = max over lanes of total steps
tot=[sum(l[ph].sum() for ph in range(3)) for l in lanes]
print('free-running bound: max lane total / mean lane total = %.3f'%(max(tot)/np.mean(tot)))
