#!/usr/bin/env python
"""numpy model of the shared-memory wavefronts of the sieve's sparse batches (bank = word & 31, one wavefront per maximum bank
multiplicity of a warp instruction): reproduces the measured 3.45 wavefronts per sparse atomic and was used to predict the
cofactor-filter experiments of profiles/r01/experiments/.  usage: tools/sim_sieve_conflicts.py [first filtered prime]"""
import numpy as np, math, sys
T=229376
N=22801763490
R=np.array([1,7,11,13,17,19,23,29])
def primes_upto(n):
    s=np.ones(n+1,bool); s[:2]=False
    for i in range(2,int(n**.5)+1):
        if s[i]: s[i*i::i]=False
    return np.nonzero(s)[0]
P=primes_upto(int(N**.5)+1)
QL=int(N//7000)+100
qs=np.zeros(QL+1,bool); qs[primes_upto(QL)]=True
base=P[P>=61]
warp_lim=T//104
sparse=base[base>=warp_lim]
PSEMI=int(sys.argv[1]) if len(sys.argv)>1 else 7646
def maxmult(banks,active):
    # banks: (n,32) ints, active (n,32) bool -> per-row max multiplicity
    n=banks.shape[0]
    cnt=np.zeros((n,32),int)
    rows=np.repeat(np.arange(n),32).reshape(n,32)
    np.add.at(cnt,(rows[active],banks[active]),1)
    return cnt.max(1)
tot_cur=0;tot_new=0;tot_instr=0;tot_gather=0; tot_new_noq=0
tiles=[100,1000,2000,3000]
for t in tiles:
    B0=t*T
    tile_end=(B0+T)*30
    act=sparse[sparse.astype(np.int64)**2<tile_end]
    for b in range(0,len(act),32):
        pb=act[b:b+32].astype(np.int64)
        n_on=len(pb)
        p=np.zeros(32,np.int64); p[:n_on]=pb
        on=np.arange(32)<n_on
        plo,phi=pb[0],pb[-1]
        U=T//phi; n_uni=U-1 if U>1 else 0
        n_it=(T-1)//plo+1
        r=np.where(on,B0%np.maximum(p,1),0)
        a0=np.where(on,B0//np.maximum(p,1),0)
        semi = plo>=PSEMI
        for j in range(8):
            c=(p*R[j])//30
            dfirst=np.where(c>=r,0,1)
            x=np.where(c>=r,c-r,c+p-r)
            if j==0: 
                sp=(B0<=c); x=x+np.where(sp,p,0); dfirst=dfirst+np.where(sp,1,0)
            ks=np.arange(n_it)[:,None]
            X=x[None,:]+ks*p[None,:]
            inT=(X<T)&on[None,:]
            banks_cur=np.where(inT,(X>>2)&31,np.arange(32)[None,:])
            w_cur=maxmult(banks_cur,np.ones_like(inT))
            tot_cur+=w_cur.sum(); tot_instr+=n_it
            if semi:
                q=30*(a0[None,:]+dfirst[None,:]+ks)+R[j]
                qok=qs[np.minimum(q,QL)]
                actv=inT&qok
            else:
                actv=inT
            w_new=maxmult((X>>2)&31,actv)
            tot_new+=w_new.sum()
            tot_new_noq+=maxmult((X>>2)&31,inT).sum()
        if semi:
            g0=(a0>>5)*32
            for off in (0,16,32,48):
                tot_gather+=len(np.unique((g0[on]+off)>>7))
nt=len(tiles)
print("per tile: instr",tot_instr/nt,"wf cur",tot_cur/nt,"wf predicated-no-q",tot_new_noq/nt,"wf new",tot_new/nt,"gather wf",tot_gather/nt)

# ---- ffs-loop variant: per (batch, strand) each lane walks its set bits; iteration i = i-th set bit of each lane
def sim_ffs(PSEMI):
    it_tot=0; wf_tot=0; inst_unf=0; wf_unf=0
    for t in tiles:
        B0=t*T
        tile_end=(B0+T)*30
        act=sparse[sparse.astype(np.int64)**2<tile_end]
        for b in range(0,len(act),32):
            pb=act[b:b+32].astype(np.int64); n_on=len(pb)
            p=np.zeros(32,np.int64); p[:n_on]=pb
            on=np.arange(32)<n_on
            plo,phi=pb[0],pb[-1]
            n_it=(T-1)//plo+1
            r=np.where(on,B0%np.maximum(p,1),0); a0=np.where(on,B0//np.maximum(p,1),0)
            semi = plo>=PSEMI and n_it<=31
            for j in range(8):
                c=(p*R[j])//30
                dfirst=np.where(c>=r,0,1)
                x=np.where(c>=r,c-r,c+p-r)
                if j==0:
                    sp=(B0<=c); x=x+np.where(sp,p,0); dfirst=dfirst+np.where(sp,1,0)
                ks=np.arange(n_it)[:,None]
                X=x[None,:]+ks*p[None,:]
                inT=(X<T)&on[None,:]
                if not semi:
                    banks_cur=np.where(inT,(X>>2)&31,np.arange(32)[None,:])
                    wf_unf+=maxmult(banks_cur,np.ones_like(inT)).sum(); inst_unf+=n_it
                    continue
                q=30*(a0[None,:]+dfirst[None,:]+ks)+R[j]
                actv=inT&qs[np.minimum(q,QL)]
                cnt=actv.sum(0)
                m=cnt.max()
                it_tot+=m
                # i-th hit per lane
                order=np.cumsum(actv,0)
                for i in range(1,m+1):
                    sel=actv&(order==i)
                    lanes=np.nonzero(sel.any(0))[0]
                    rows=sel.argmax(0)[lanes]
                    bk=(X[rows,lanes]>>2)&31
                    wf_tot+=np.bincount(bk,minlength=32).max()
    nt=len(tiles)
    print("PSEMI",PSEMI,"filtered: iterations",it_tot/nt,"wf",wf_tot/nt,"| unfiltered rest: instr",inst_unf/nt,"wf",wf_unf/nt)
for ps in (7400,):
    sim_ffs(ps)
