"""Development aid (CPU): factor/solve trips of the constraint solver per substep, per lane and per warp (max over 32\nconsecutive particles), from the product device math compiled for the host (tests/hostcheck).  The warp-level trip\ncount is what the GPU pays: one lane that needs a third pass keeps its whole warp in the loop.\nusage: python tools/host_trip_stats.py [-DEFINE ...]"""
import os, sys, subprocess, ctypes as C, numpy as np
ROOT='/root/repo'; sys.path.insert(0, ROOT); sys.path.insert(0, ROOT+'/tests')
from conftest import reference_noise, synthetic_state
from mjmpc_b200.envs.model import compile_model, reacher7dof_spec
cm = compile_model(reacher7dof_spec()); P = cm.chain.params
def _p(a): return a.ctypes.data_as(C.POINTER(C.c_double))
defs=[d for d in sys.argv[1:]]
so='/tmp/libtrips_%s.so'%("_".join(defs) or "default")
subprocess.check_call(["g++","-std=c++17","-O2","-fPIC","-shared","-o",so,ROOT+"/tests/hostcheck/hostcheck.cpp","-lm"]+["-D"+d for d in defs])
lib=C.CDLL(so)
K,H=2048,32
tot=np.zeros(6); 
for seed in range(4):
    st=synthetic_state(cm,seed); noise=reference_noise(K,H,7,seed+10); mean=np.zeros((H,7))
    trips=np.zeros(K*H*2,dtype=np.int32)
    lib.hostcheck_record_trips(trips.ctypes.data_as(C.POINTER(C.c_int)))
    costs=np.zeros((K,H)); qv=np.zeros((K,H,14))
    lib.hostcheck_rollout(_p(P),0,_p(st["qp"]),_p(st["qv"]),_p(st["target_pos"]),K,H,_p(mean),_p(noise),_p(costs),_p(qv))
    lib.hostcheck_record_trips(None)
    t=trips.reshape(K,H*2)          # particle-major, substep order
    w=t.reshape(K//32,32,H*2).max(axis=1)      # warp-level trips per substep
    tot+= [ (t>1).mean(), t[t>1].mean(), (t>=3).mean(), w.mean(), (w>=3).mean(), (w>=4).mean()]
tot/=4
print("%-24s constrained %.3f  trips/constrained %.3f  lanes>=3 trips %.4f | warp trips mean %.3f  warps>=3 %.3f  warps>=4 %.3f"%((",".join(defs) or "default",)+tuple(tot)))
