"""Developer script (CPU, needs /root/reference): per-band GP host instantiation (gp1d.cuh + lbfgsb.cuh) against the
reference's fit_gp_single_band (scikit-learn) on small and odd bands: 5-13 points, pure noise, a single bump, a ramp,
quantised fluxes, nearly constant fluxes; spans of 5 / 60 / 800 days; errors of 0.01 / 0.3 / 3.
    python tests/fuzz_gp1d_vs_reference.py <seed> <objects>      (4 bands per object, ~0.1 s per band upstream)"""
import ctypes, os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, '/root/reference/src')
import numpy as np
from packref import pack_reference
from hostsim.build import build
from fitparity import classify_gp1d
from features.gaussian_process import fit_gp_single_band
warnings.simplefilter('ignore')
lib = ctypes.CDLL(build()); P=ctypes.c_void_p
rng=np.random.default_rng(int(sys.argv[1]))
n_obj=int(sys.argv[2])
obj,band,t,f,e=[],[],[],[],[]
for k in range(n_obj):
    for b in (1,2,3,4):
        n=int(rng.integers(5,14))
        mode=rng.integers(0,5)
        tt=np.sort(60000+rng.uniform(0,rng.choice([5.0,60.0,800.0]),n))
        if mode==0: ff=rng.normal(0,1,n)
        elif mode==1: ff=10*np.exp(-0.5*((tt-tt.mean())/ (0.2*(tt[-1]-tt[0]+1)))**2)+rng.normal(0,0.3,n)
        elif mode==2: ff=np.linspace(1,5,n)+rng.normal(0,0.05,n)
        elif mode==3: ff=rng.integers(0,3,n).astype(float)
        else: ff=rng.normal(5,0.01,n)
        ee=rng.choice([0.01,0.3,3.0])*rng.uniform(0.5,1.5,n)
        obj+=[k]*n; band+=[b]*n; t+=list(tt); f+=list(ff); e+=list(ee)
obj=np.array(obj,np.int32); band=np.array(band,np.uint8); t=np.array(t); f=np.array(f); e=np.array(e)
pk=pack_reference(obj,band,t,f,e,n_obj)
out=np.zeros((n_obj,21)); st=np.zeros((n_obj,12),np.int32)
args = [ctypes.c_longlong(n_obj)] + [pk[k].ctypes.data_as(P) for k in ("seg_off","t","flux","err","rank")] + [out.ctypes.data_as(P), st.ctypes.data_as(P)]
lib.hs_gp1d(*args)
ref=np.full((n_obj,21),np.nan)
for k in range(n_obj):
    for j,b in enumerate((1,2,3,4)):
        m=(obj==k)&(band==b)
        o=np.argsort(t[m],kind='stable')
        _,feat=fit_gp_single_band(t[m][o],f[m][o],e[m][o])
        ref[k,4*j:4*j+4]=[feat['gp_length_scale'],feat['gp_amplitude'],feat['gp_noise'],feat['gp_log_likelihood']]
c=classify_gp1d(out[:, :16].copy() if False else np.hstack([out[:,:16],np.full((n_obj,5),np.nan)]), np.hstack([ref[:,:16],np.full((n_obj,5),np.nan)]))
print(c)
for k in range(n_obj):
    for j in range(4):
        r,m_=ref[k,4*j:4*j+4],out[k,4*j:4*j+4]
        if not np.allclose(m_,r,rtol=1e-3,atol=1e-6) and not (m_[3] >= r[3]-1e-6*abs(r[3])-1e-9):
            print("WORSE obj",k,"band",j,"ours",m_,"ref",r,"status",st[k,3*j:3*j+3])
