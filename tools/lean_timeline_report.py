"""Reads gpurun_out/lean_trace.bin (tools/lean_timeline_patch.py): when the CTAs of one k_weight_lean launch end, how long warps idle
before their CTA ends, chunk durations.  usage: lean_timeline_report.py [file]"""
import sys

import numpy as np

for name in [sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/lean_trace.bin']:
    t=np.fromfile(name,dtype=np.uint64).reshape(148,32,16).astype(np.int64)
    t0=t[t>0].min()
    ns=(t>0).sum(axis=2)
    T=(t-t0)/1e3
    last=np.array([[T[c,w,ns[c,w]-1] for w in range(32)] for c in range(148)])
    lastchunk=np.array([[T[c,w,max(ns[c,w]-2,0)] for w in range(32)] for c in range(148)])
    print("==",name,"CTA end us: min/mean/max %.1f %.1f %.1f"%(last.max(1).min(), last.max(1).mean(), last.max(1).max()), " per-warp idle before CTA end mean %.1f"%((last.max(1)[:,None]-lastchunk).mean()))
    d=[];firsts=[]
    nchunks=0
    for c in range(148):
        for w in range(32):
            k=ns[c,w]
            ts=T[c,w,:k-1]
            dd=np.diff(ts)
            if len(dd)>0: firsts.append(dd[0])
            d.extend(dd[1:]); nchunks+=len(dd)
    d=np.array(d); firsts=np.array(firsts)
    print("  chunks",nchunks,"steady chunk us p10/p50/p90/max", np.round(np.percentile(d,[10,50,90]),1), round(d.max(),1), " first chunk mean %.1f"%firsts.mean(), " n>40us:", (d>40).sum())
    print("  chunks per warp hist", np.bincount(ns.ravel()-2))
    print("  slots saturated (>=16):", (ns>=16).sum())
    e=np.sort(last.max(1)); print("  CTA ends:", np.round(e[::12],1))
    # per-CTA total chunks
    cc=(ns-2).sum(1); print("  chunks per CTA min/mean/max", cc.min(), cc.mean(), cc.max())
