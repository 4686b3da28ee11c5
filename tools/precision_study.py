"""CPU (numpy, fp64) emulation of the fp16 rounding strategies considered for the tensor-core sweep: which operand
rounding dominates the error of q = k^T K^-1 k (sigma^2 = a - q) and of grad sigma, per data distribution.
Output committed as profiles/r01_tensor_precision_study.txt. Not used by the product."""
import sys, numpy as np
sys.path.insert(0,'/root/repo/tools')
import synth
def study(N,D,M,kind):
    X=synth.make_X(N,D,kind); th=synth.make_theta(D,'default'); a=th[0]; l=th[1:]; b=0.005
    Q=synth.make_queries(M,D)
    def kern(A,B):
        d=(A[:,:,None]-B[:,None,:])/l[:,None,None]
        return a*np.exp(-0.5*(d**2).sum(0))
    K=kern(X,X)+b*np.eye(N); Ki=np.linalg.inv(K)
    L=np.linalg.cholesky(K); W=np.linalg.inv(L)
    ks=kern(X,Q)
    def f16(v,s): return (v*s).astype(np.float16).astype(np.float64)/s
    sc=lambda v: 2**np.floor(np.log2(32768/np.abs(v).max()))
    kt=f16(ks,sc(ks)); At=f16(Ki,sc(Ki)); Wt=f16(W,sc(W))
    u=Ki@ks; q=(ks*u).sum(0); sig=np.sqrt(a-q)
    # gradient pieces: P2_d = sum_i X_di k_i u_i ; dsig_d = -(1/sig)*(-2)*(x_d q - P2_d)/l^2  (c=2)
    def dsig(qq,P2,ss): return -(1/ss)*(-2)*(Q*qq-P2)/(l[:,None]**2)
    P2=X@(ks*u); ds=dsig(q,P2,sig)
    def rep(name,qq,P2x):
        ss=np.sqrt(np.maximum(a-qq,1e-300)); dd=dsig(qq,P2x,ss)
        print(f"  {name:26s} sigma err max {np.abs(ss-sig).max()/sig.max():.2e} rms {np.sqrt(((ss-sig)**2).mean())/sig.max():.2e} | dsigma err max {np.abs(dd-ds).max()/np.abs(ds).max():.2e} rms {np.sqrt(((dd-ds)**2).mean())/np.abs(ds).max():.2e}")
    print(f"N={N} D={D} kind={kind} cond {np.linalg.cond(K):.1e}")
    ut=At@kt; rep('Kinv 1-pass', (kt*ut).sum(0), X@(kt*ut))
    Alo=f16(Ki-At,sc(Ki)); u2=(At+Alo)@kt
    rep('Kinv hi+lo + kcorr(q)', ((kt+2*(ks-kt))*u2).sum(0), X@(kt*u2))
    rep('Kinv hi+lo + kcorr(q,P2~)', ((kt+2*(ks-kt))*u2).sum(0), X@(ks*u2))
    v=Wt@kt; qv=(v*v).sum(0); v16=f16(v,sc(v)); beta=Wt.T@v16
    rep('W tri: q=|v|^2, b=W^T v16', qv, X@(kt*beta))
    # with k correction on v: v = W k~ + W dk -> q ~ |v|^2 + 2 v.(W dk) needs another GEMM: skip
    Wlo=f16(W-Wt,sc(W)); v2=(Wt+Wlo)@kt; q2=(v2*v2).sum(0); 
    rep('W hi+lo tri (q only)', q2, X@(kt*beta))
study(96,6,200,'uniform'); study(700,16,2000,'sls'); study(2048,16,1500,'uniform'); study(2048,16,1000,'sls')
