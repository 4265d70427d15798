import sys, os, time, ctypes as C
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from inverted_index_2_b200 import synth, _abi as A
from inverted_index_2_b200.engine import Engine
from inverted_index_2_b200.flat import FlatSegment, views_array
eng = Engine(0)
w = synth.make_workload(1000000, 64, 100000000, seed=0xC2, removed_frac=0.05)
def pin(a):
    a = np.ascontiguousarray(a)
    t = torch.from_numpy(a).pin_memory()
    return t.numpy(), t
keep=[]; hsegs=[]
for s in w.segments:
    arrs={}
    for f in ("term_bytes","term_off","post","post_off"):
        arrs[f],k=pin(getattr(s,f)); keep.append(k)
    hsegs.append(FlatSegment(arrs["term_bytes"],arrs["term_off"],s.mode,post=arrs["post"],post_off=arrs["post_off"]))
hrem,k=pin(w.removed); keep.append(k)
arr=views_array(hsegs); out=A.MergeOut()
def step():
    eng._check(eng.lib.ii2_merge(arr,len(hsegs),A.np_ptr(hrem,A.u32p),len(hrem),0,C.byref(out)),"merge")
    eng.lib.ii2_merge_out_free(C.byref(out))
for P in sys.argv[1:]:
    os.environ["II2_MERGE_PARTS"]=P
    for _ in range(3): step()
    eng.prof_enable(True)
    t0=time.perf_counter()
    for _ in range(5): step()
    dt=(time.perf_counter()-t0)/5
    pr=eng.prof_read(); eng.prof_enable(False)
    print("P",P,"ms",round(dt*1e3,2)," ".join("%s=%.2f/%.2f(x%d)"%(k['name'][:18],k['ms']/5,k['host_ms']/5,k['count']//5) for k in pr))
