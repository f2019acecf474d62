"""Developer script (CPU): AddressSanitizer + UBSan over the host instantiation of every device template (the same
source the CUDA kernels instantiate: stat / colors / shape / tde / physics / research rows, the thread-per-fit TRF jobs,
the GP host model, the per-band GP with both objectives) on the edge / sparse / dense fixtures, the crafted per-band GP
cases and fuzzed degenerate objects.  compute-sanitizer is not available on the GPU pool; this covers the indexing and
arithmetic of the shared templates (not the GPU-only parallel code in stat_fast.cu, gp2d.cu, pack.cu).

    g++ -O1 -g -std=c++17 -fPIC -shared -fsanitize=address,undefined -fno-omit-frame-pointer -x c++ \\
        -I mallorn_astrophysics_b200/csrc tests/hostsim/hostsim.cpp -o /tmp/libhostsim_asan.so -lm
    LD_PRELOAD=$(g++ -print-file-name=libasan.so):$(g++ -print-file-name=libubsan.so) ASAN_OPTIONS=detect_leaks=0 \\
        python tests/hostsim_asan.py
"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from packref import pack_reference
from fuzz_hostsim_vs_reference import make_fuzz
lib = ctypes.CDLL('/tmp/libhostsim_asan.so'); P=ctypes.c_void_p
def run_all(obj,band,t,f,e,z,n_obj,tag):
    pk=pack_reference(obj,band,t,f,e,n_obj)
    base=[ctypes.c_longlong(n_obj)]+[pk[k].ctypes.data_as(P) for k in ("seg_off","t","flux","err","rank")]
    zz=np.ascontiguousarray(z)
    for fn,nc,takes_z,status in (("hs_stat",123,False,0),("hs_color",83,True,0),("hs_shape",65,True,0),("hs_tde",25,True,0),("hs_physics",32,True,0),
                                 ("hs_research",40,True,0),("hs_bazin_t",52,False,12),("hs_powerlaw_t",27,False,54),("hs_gp2d",27,False,4),("hs_gp1d",21,False,12)):
        out=np.zeros((n_obj,nc)); args=list(base)
        if takes_z: args.append(zz.ctypes.data_as(P))
        args.append(out.ctypes.data_as(P))
        if status:
            st=np.zeros((n_obj,status),np.int32); args.append(st.ctypes.data_as(P))
        getattr(lib,fn)(*args)
    print(tag,"ok",n_obj,"objects")
for name in ("ref_edge","ref_sparse","ref_dense"):
    g=np.load(os.path.join(ROOT, 'tests', 'golden', f'{name}.npz'))
    run_all(g["obj"],g["band"],g["t"],g["flux"],g["err"],g["z"],int(g["n_obj"]),name)
g=np.load(os.path.join(ROOT, 'tests', 'golden', 'gp1d_crafted.npz'))
run_all(g["obj"],g["band"],g["t"],g["flux"],g["err"],np.full(10,0.3),10,"crafted")
obj,band,t,f,e,z=make_fuzz(120,3)
run_all(obj,band,t,f,e,z,120,"fuzz")
