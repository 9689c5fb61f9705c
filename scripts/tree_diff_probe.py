import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from quickrank_b200 import api, synth
from oracle import pyref
import bench
w=bench.WORKLOAD
x,l,off=synth.make_dataset(w["n_docs"], w["n_features"], w["n_queries"], seed=w["seed"])
ref=bench.oracle_first_trees(x,l,off,3)
tr=api.Trainer(x,l,off,algo="LAMBDAMART",nleaves=64,minleafsupport=1,cutoff=10,shrinkage=0.1,hist_mode=api.HIST_FAST)
for m in range(3):
    t,_=tr.boost_iteration(want_tree=True)
    r=ref[m]
    n=min(len(t["feature"]),len(r["feature"]))
    d=[i for i in range(n) if t["feature"][i]!=r["feature"][i] or t["threshold_idx"][i]!=r["threshold_idx"][i]]
    print("tree",m,"nodes",len(t["feature"]),len(r["feature"]),"first diffs",d[:5])
    for i in d[:3]:
        print("  node",i,"gpu f,t,count",t["feature"][i],t["threshold_idx"][i],t["count"][i],"ref",r["feature"][i],r["threshold_idx"][i],r["count"][i], "dev", t["deviance"][i], r["deviance"][i])
