import sys, numpy as np
sys.path.insert(0,'/root/repo')
from inverted_index_2_b200 import synth
from inverted_index_2_b200.engine import Engine
eng=Engine(0)
w = synth.make_workload(1000000, 256, 100000000, seed=0xC3, presence=0.125)
dsegs=[eng.upload(s) for s in w.segments]; drem=eng.upload_removed(w.removed)
n=1000000
for span in (1000, 10000):
    lo=500000
    tlo=synth.term_at(w.term_bytes,w.term_off,lo); thi=synth.term_at(w.term_bytes,w.term_off,lo+span-1)
    for i in range(3):
        r=eng.read_range_dev(dsegs,tlo,thi,drem); r.release()
    eng.prof_enable(True)
    for i in range(5):
        r=eng.read_range_dev(dsegs,tlo,thi,drem); r.release()
    print(span, [(k['name'], round(k['ms']/k['count'],3), round(k['host_ms']/k['count'],3)) for k in eng.prof_read()])
    eng.prof_enable(False)
