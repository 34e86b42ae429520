"""In-container check (needs /root/reference, so it is not a test): the CPU restatement of the interior-point solve on
ALL valid stored IPOPT solutions of the CCC problem (optimizations/landing/data/*.mat; the test suite uses the 43 of them
committed as ccc_n41.npz).  Round-1 result:
    stored runs 294, converged 286, same local solution as IPOPT 190 (cost within 0.2 %, identical touchdown knots;
    terminal state within 9.6e-4, median max|df_z| 0.32 N), lower cost than IPOPT's 65, higher cost 31, mean 86.5 iterations
Run:  python tests/golden/check_all_stored.py"""
import sys, glob
import os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import numpy as np, scipy.io as sio, ccc_problem as ccc
from oracle_ip import solve_cpu, default_options, default_problem
X=[];F=[];TD=[]
for path in sorted(glob.glob('/root/reference/optimizations/landing/data/*.mat')):
    for s in np.atleast_1d(sio.loadmat(path, squeeze_me=True, struct_as_record=False)['opt_sol']):
        if s.X_star.shape==(12,41) and abs(s.X_star[2,0]-0.6)<1e-9:
            X.append(s.X_star); F.append(s.f_star); TD.append(np.asarray(s.td,float))
X=np.array(X);F=np.array(F);TD=np.array(TD)
drops=np.ascontiguousarray(X[:,:,0])
opt=default_options(run_Qf=list(ccc.QF), kin_box=list(ccc.KIN_BOX))
r=solve_cpu(ccc.N, drops, opt=opt, pb=ccc.fill_problem(default_problem()))
same=0; lower=0; higher=0; dfz=[]; dterm=[]
for b in range(len(drops)):
    if r['status'][b]!=0: continue
    Xs,cs,fs=ccc.split(r['x'][b]); fref=ccc.stored_cost(X[b],F[b])
    rel=(r['f'][b]-fref)/fref
    if abs(rel)<=2e-3 and ccc.touchdown(fs)==TD[b].astype(int).tolist():
        same+=1; dfz.append(np.max(np.abs(fs[2::3]-F[b][2::3]))); dterm.append(max(np.max(np.abs(Xs[2:5,-1]-X[b][2:5,-1])), np.max(np.abs(Xs[6:,-1]-X[b][6:,-1]))))
    elif rel<0: lower+=1
    else: higher+=1
print("stored runs", len(drops), "converged", int((r['status']==0).sum()), "same", same, "lower cost", lower, "higher cost", higher, "median dfz %.2f max dterm %.2e"%(np.median(dfz), max(dterm)), "mean iters %.1f"%r['iters'].mean())
