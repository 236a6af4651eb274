"""numpy emulation of the join passes (logic only): permuted sort on both sides, equal-key buckets, a = min, first condition,
exact D and S from the oracle -> must equal the oracle's edge set."""
import sys, ctypes as C, subprocess, os, numpy as np
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'tests'))
from oracle import oracle as orc
from badger_b200 import synth
import tempfile
so=os.path.join(tempfile.mkdtemp(prefix='seed2_'), 'shim.so')
subprocess.check_call(["g++","-O2","-std=c++17","-shared","-fPIC","-x","c++",os.path.join(ROOT,"tests","core_host_shim.cpp"),"-o",so])
L=C.CDLL(so)
u32p=np.ctypeslib.ndpointer(np.uint32, flags="C")
L.shim_seed2_permute.argtypes=[C.c_int,u32p,u32p,C.c_size_t,u32p,u32p,u32p,u32p]
L.shim_seed2_first.argtypes=[u32p,u32p,C.c_size_t,np.ctypeslib.ndpointer(np.int8, flags="C")]
L.shim_seed2_key_bits.restype=C.c_int
rng=synth.rng_for(91)
cells=rng.integers(0,1<<32,300,dtype=np.uint64).astype(np.uint32)
obs,_=synth.simulate_reads(cells,30000,0.06,rng)
s=np.unique(obs); n=s.size
wa,wb,wd,_=orc.Index(s).edges(2)
want=set(zip(wa.tolist(),wb.tolist(),wd.tolist()))
OL=orc.lib()
got=set(); cand=0
pa=np.zeros(n,np.uint32); pb=np.zeros(n,np.uint32); ua=np.zeros(n,np.uint32); ub=np.zeros(n,np.uint32)
for c in range(20):
    L.shim_seed2_permute(c,s,s,n,pa,pb,ua,ub)
    bits=L.shim_seed2_key_bits(c)
    sa=np.sort(pa); sb=np.sort(pb)
    ka=sa>>np.uint32(32-bits); kb=sb>>np.uint32(32-bits)
    lo=np.searchsorted(kb,ka,'left'); hi=np.searchsorted(kb,ka,'right')
    L.shim_seed2_permute(c,s,s,n,pa,pb,ua,ub)   # (unpermute check happens in the unit test)
    # recover barcodes of the sorted words
    # rows
    xs=np.zeros(n,np.uint32); ys=np.zeros(n,np.uint32)
    tmp=np.zeros(n,np.uint32)
    # unpermute sorted words by permuting lookup: build dict from permuted -> original
    mapa=dict(zip(pa.tolist(),s.tolist())); mapb=dict(zip(pb.tolist(),s.tolist()))
    for i in range(n):
        if hi[i]>lo[i]:
            x=mapa[int(sa[i])]
            for j in range(lo[i],hi[i]):
                y=mapb[int(sb[j])]
                if x<y:
                    cand+=1
                    first=np.zeros(1,np.int8); L.shim_seed2_first(np.asarray([x],np.uint32),np.asarray([y],np.uint32),1,first)
                    if int(first[0])==c:
                        d=OL.orc_D(x,y)
                        if d<=2 and OL.orc_S(x,y)>=4: 
                            assert (x,y,d) not in got
                            got.add((x,y,d))
print("N",n,"oracle edges",len(want),"join edges",len(got),"equal",got==want,"candidates",cand,"of",n*(n-1)//2)
