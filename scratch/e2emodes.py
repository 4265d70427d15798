"""e2e staging experiment: ii2_merge over pinned host buffers under every II2_MERGE_UPLOAD mode.
usage: python scratch/e2emodes.py "<MODE>[:PARTS[:GRID]]" ...   (MODE '-' = default gather)"""
import sys, os, time, zlib, ctypes as C
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
def step(check=False):
    eng._check(eng.lib.ii2_merge(arr,len(hsegs),A.np_ptr(hrem,A.u32p),len(hrem),0,C.byref(out)),"merge")
    sig=None
    if check:
        T=out.terms_count
        tb=np.ctypeslib.as_array(out.term_off,(T+1,))
        sig=(T,out.val_size,
             zlib.crc32(np.ctypeslib.as_array(out.term_bytes,(int(tb[T]),)).tobytes()),
             zlib.crc32(tb.tobytes()),
             zlib.crc32(np.ctypeslib.as_array(out.val_off,(T,)).tobytes()),
             zlib.crc32(np.ctypeslib.as_array(out.val_bytes,(out.val_size,)).tobytes()))
    eng.lib.ii2_merge_out_free(C.byref(out))
    return sig
base=None
for spec in sys.argv[1:]:
    f=spec.split(":")
    for k in ("II2_MERGE_UPLOAD","II2_MERGE_PARTS","II2_GATHER_GRID"): os.environ.pop(k,None)
    if f[0]!="-": os.environ["II2_MERGE_UPLOAD"]=f[0]
    if len(f)>1 and f[1]: os.environ["II2_MERGE_PARTS"]=f[1]
    if len(f)>2 and f[2]: os.environ["II2_GATHER_GRID"]=f[2]
    sig=step(True)
    if base is None: base=sig
    for _ in range(2): step()
    ts=[]
    for _ in range(7):
        t0=time.perf_counter(); step(); ts.append(time.perf_counter()-t0)
    ts.sort()
    print("%-16s median %.2f ms  min %.2f  max %.2f  same_bytes=%s"%(spec,ts[3]*1e3,ts[0]*1e3,ts[-1]*1e3,sig==base),flush=True)
