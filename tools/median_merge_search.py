"""Search behind a DESIGN.md section 9 remark: how many comparators does it take to select the median of 25 once the five
columns of the 5x5 window are sorted (column sorts can be shared by the five horizontally neighbouring output pixels)?
Batcher merges of the sorted columns, dead-comparator elimination towards output rank 13, then greedy removal verified on ALL
7776 zero-one inputs with sorted columns (zero-one principle).  Result: 66 comparators (+ 9 per shared column sort) against the
99 of the full network in pf_math.cuh.  CPU only:  python tools/median_merge_search.py"""
import itertools, random, sys
import numpy as np
pats=list(itertools.product(range(6),repeat=5))
N=len(pats)
W0=np.zeros((25,N),dtype=bool)
for i,p in enumerate(pats):
    for c,k in enumerate(p):
        for r in range(5):
            W0[c*5+r,i]= r>=5-k
want=np.array([sum(p)>=13 for p in pats])

def oe_merge_pow2(lo, n, r, out):
    # Batcher odd-even merge on wires lo..lo+n-1 (n power of 2), stride r
    step = r*2
    if step < n:
        oe_merge_pow2(lo, n, step, out)
        oe_merge_pow2(lo+r, n, step, out)
        for i in range(lo+r, lo+n-r, step):
            out.append((i, i+r))
    else:
        out.append((lo, lo+r))

def merge_lists(a, b):
    """a, b: lists of wire ids, each sorted ascending along the list. returns (comparators, merged wire list sorted)"""
    n=1
    while n < max(len(a),len(b)): n*=2
    # virtual wires: first half a padded with +inf (None) at the end, second half b padded
    va=a+[None]*(n-len(a)); vb=b+[None]*(n-len(b))
    virt=va+vb
    comps=[]
    oe_merge_pow2(0, 2*n, 1, comps)
    real=[]
    # simulate: None = +inf stays at top; a comparator (i,j) with j None: no-op; with i None and j real: swap virtual positions
    pos=list(virt)
    for i,j in comps:
        x,y=pos[i],pos[j]
        if x is None and y is None: continue
        if y is None: continue
        if x is None:
            pos[i],pos[j]=y,None   # real goes down, inf goes up: wire relabel, no comparator needed
            continue
        real.append((x,y))
    merged=[p for p in pos if p is not None]
    assert len(merged)==len(a)+len(b)
    return real, merged

def run(net,out_wire):
    v=[W0[w].copy() for w in range(25)]
    for a,b in net:
        lo=v[a]&v[b]; hi=v[a]|v[b]; v[a]=lo; v[b]=hi
    return v[out_wire]

def build(order):
    cols=[[c*5+r for r in range(5)] for c in order]
    net=[]
    c1,m1=merge_lists(cols[0],cols[1]); net+=c1
    c2,m2=merge_lists(cols[2],cols[3]); net+=c2
    c3,m3=merge_lists(m1,m2); net+=c3
    c4,m4=merge_lists(m3,cols[4]); net+=c4
    return net, m4[12]

def liveness(net,out):
    live={out}; keep=[]
    for a,b in reversed(net):
        if a in live or b in live:
            keep.append((a,b)); live.add(a); live.add(b)
    return keep[::-1]

def prune(net,out):
    net=list(net)
    changed=True
    while changed:
        changed=False
        for idx in range(len(net)-1,-1,-1):
            trial=net[:idx]+net[idx+1:]
            if np.array_equal(run(trial,out),want):
                net=trial; changed=True
    return net
net,out=build([0,1,2,3,4])
assert np.array_equal(run(net,out),want)
print('full',len(net))
net=liveness(net,out); print('live',len(net))
net=prune(net,out); print('pruned',len(net))
print(net,out)
