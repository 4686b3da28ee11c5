"""CPU (numpy, fp64) emulation of the tensor sweep's split-fp16 modes: where does the gradient error of the 2-pass mode
(SLSGP_SWEEP_TENSOR_X2: k16.A_hi + k16.A_lo, first-order correction of q in the epilogue) sit, and would a second tier keyed on
sigma^2 / a remove it? For each data distribution: max error of sigma and grad sigma (relative to the largest reference entry) over
all candidates, and over the candidates that a second tier with threshold tau would leave on the tensor path.
Output committed as profiles/r02t_x2_error_study.txt. Not used by the product."""
import sys, numpy as np
sys.path.insert(0,'/root/repo/tools')
import synth
def f16(v,s): return (v*s).astype(np.float16).astype(np.float64)/s
def sc(v): return 2.0**np.floor(np.log2(32768/np.abs(v).max()))
def study(N,D,M,kind,thetakind='default',near=0):
    X=synth.make_X(N,D,kind); th=synth.make_theta(D,thetakind); a=th[0]; l=th[1:]; b=0.005
    Q=synth.make_queries(M,D)
    if near: Q[:, :near] = X[:, :near] + 0.02*np.random.default_rng(5).standard_normal((D,near))
    def kern(A,B):
        d=(A[:,:,None]-B[:,None,:])/l[:,None,None]
        return a*np.exp(-0.5*(d**2).sum(0))
    K=kern(X,X)+b*np.eye(N); Ki=np.linalg.inv(K)
    y=synth.make_y(X); alpha=Ki@y
    ks=kern(X,Q)
    sK=2.0**np.floor(np.log2(32768/a)); sA=sc(np.diag(Ki))
    kt=f16(ks,sK); dk=f16(ks-kt,sK); Ah=f16(Ki,sA); Al=f16(Ki-Ah,sA)
    c=2.0
    def outputs(u,kw_q,kw_g):
        q=(kw_q*u).sum(0); sig2=a-q; sig=np.sqrt(np.maximum(sig2,1e-300))
        gb=-c*q; P2=-c*(X@(kw_g*u))
        ds=-(1/sig)*((Q*gb-P2)/(l[:,None]**2))
        return sig2,sig,ds
    u=Ki@ks; s2,sg,ds=outputs(u,ks,ks)
    res={}
    u3=(Ah+Al)@kt+Ah@dk; res['3pass']=outputs(u3,kt+dk,kt+dk)
    u2=(Ah+Al)@kt; res['x2']=outputs(u2,kt+2*dk,kt+dk)
    u1=Ah@kt; res['x1']=outputs(u1,kt,kt)
    # x2 variant: second pass is k_lo*A_hi instead (drop A_lo)
    print(f"N={N} D={D} {kind} theta={thetakind}: sigma2/a quantiles {np.quantile(s2/a,[0,.01,.1,.5])}")
    mds=np.abs(ds).max()
    for name,(s2x,sgx,dsx) in res.items():
        e_sig=np.abs(sgx-sg)/sg.max(); e_ds=np.abs(dsx-ds).max(0)/mds
        line=f"  {name:6s} all: sig {e_sig.max():.2e} dsig {e_ds.max():.2e}"
        for tau in (0.1,0.2,0.3,0.4,0.5):
            keep=s2/a>=tau
            mk=np.abs(ds[:,keep]).max() if keep.any() else 1
            line+=f" | tau {tau}: drop {100*(1-keep.mean()):.1f}% sig {e_sig[keep].max():.1e} dsig {(np.abs(dsx-ds).max(0)[keep]).max()/mds:.1e}"
        print(line)
study(2048,16,3000,'uniform')
study(2048,16,3000,'sls')
study(2048,16,3000,'sls',near=300)
study(2048,16,3000,'uniform','perturbed')
study(700,16,3000,'sls','perturbed',near=200)
study(512,8,3000,'uniform')
study(448,16,3000,'sls','perturbed')
