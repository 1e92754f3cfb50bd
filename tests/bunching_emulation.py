"""How accurate is the REFERENCE's float64 arithmetic on the heavy-bunching item of tests/test_gpu_edges.py (S = 20 + 20 particles in two
modes, T = 4 x 10)?  CPU only.  The Chin-Huh sum cancels ~7 digits there; this script evaluates it in float64 under several evaluation
orders that are all "the reference's algorithm" up to rounding (fresh or incremental sums, lexicographic or Guan order, full or
symmetry-halved walk, pow by binary exponentiation or a product tree) and compares with the oracle's 80-bit value: the results
scatter over 4e-10 .. 9e-10, while the oracle's double restatement of the exact reference order happens to land at 9e-11.  The kernel's
8e-10 on this item (tests/bunching_accuracy.py on a GPU) lies inside that scatter.

    python tests/bunching_emulation.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, math, itertools
from tests import workloads
from oracle import pyoracle as orc
U = workloads.haar(8, 40)
S = np.array([20, 20, 0, 0, 0, 0, 0, 0]); T = np.array([10, 10, 10, 10, 0, 0, 0, 0])
truth = orc.guan_permanent(U, S, T, orc.CHIN_HUH, "ld")
print("truth", truth, "ref-d err", abs(orc.guan_permanent(U, S, T, orc.CHIN_HUH, "d")-truth)/abs(truth))
walk=[v for v in range(8) if S[v]>0]; w=[int(S[v]) for v in walk]
cols=[j for j in range(8) if T[j]>0]; mult=[int(T[j]) for j in cols]
n=int(S.sum())
def cpowi(a,e):
    r=1+0j; b=a
    while e>0:
        if e&1: r=r*b
        e>>=1
        if e: b=b*b
    return r
def tree(vals):
    if len(vals)==1: return vals[0]
    mid=len(vals)//2
    return tree(vals[:mid])*tree(vals[mid:])
def run(halved, product, order="lex", flush=None):
    total=0j; terms=[]
    ranges=[range(x+1) for x in w]
    if halved: ranges[-1]=range(w[-1]//2+1)
    for r in itertools.product(*ranges):
        coef=[w[i]-2*r[i] for i in range(len(w))]
        sums=[sum(coef[i]*U[cols[jj]][walk[i]] for i in range(len(w))) for jj in range(len(cols))]   # fresh sums: U[out][in], product side = outputs
        if product=="pow":
            p=1+0j
            for jj in range(len(cols)): p=p*cpowi(sums[jj],mult[jj])
        else:
            vals=[]
            for jj in range(len(cols)): vals+= [sums[jj]]*mult[jj]
            p=tree(vals)
        b=1
        for i in range(len(w)): b*=math.comb(w[i],r[i])
        wt=1
        if halved: wt = 2 if 2*r[-1]<w[-1] else 1
        sign=-1 if sum(r)&1 else 1
        terms.append(sign*b*wt*p)
    for t in terms: total+=t
    return total/2**n
for halved in (False, True):
    for product in ("pow","tree"):
        v=run(halved, product)
        print("halved",halved,"product",product,"rel err", abs(v-truth)/abs(truth))

def guan_sequence(lims):
    # reflected mixed-radix Gray code: yields digit vectors, consecutive ones differ by +-1 in one digit (digit 0 fastest)
    D=len(lims); r=[0]*D; d=[1]*D
    yield tuple(r)
    while True:
        v=0
        while v<D:
            nxt=r[v]+d[v]
            if 0<=nxt<=lims[v]: break
            d[v]=-d[v]; v+=1
        if v==D: return
        r[v]+=d[v]
        yield tuple(r)

def run_incremental(chunk=None, halved=False):
    lims=list(w)
    if halved: lims[-1]=w[-1]//2
    seq=list(guan_sequence(lims))
    total=0j; sums=None
    for idx,r in enumerate(seq):
        if sums is None or (chunk and idx%chunk==0):
            coef=[w[i]-2*r[i] for i in range(len(w))]
            sums=[sum(coef[i]*U[cols[jj]][walk[i]] for i in range(len(w))) for jj in range(len(cols))]
        else:
            v=[i for i in range(len(w)) if r[i]!=prev[i]][0]; delta=r[v]-prev[v]
            sums=[sums[jj]-2*delta*U[cols[jj]][walk[v]] for jj in range(len(cols))]
        prev=r
        p=1+0j
        for jj in range(len(cols)): p=p*cpowi(sums[jj],mult[jj])
        b=1
        for i in range(len(w)): b*=math.comb(w[i],r[i])
        wt=1
        if halved: wt = 2 if 2*r[-1]<w[-1] else 1
        sign=-1 if sum(r)&1 else 1
        total+=sign*b*wt*p
    return total/2**n
for halved in (False, True):
    for chunk in (None, 21, 63):
        v=run_incremental(chunk, halved)
        print("incremental guan order, halved",halved,"fresh start every",chunk,"rel err", abs(v-truth)/abs(truth))
