"""CPU probe (oracle arithmetic): which fraction of the C2 workload the decision filter of
csrc/filter.cu decides with the mean alone / with the variance given the first R training rows."""
import numpy as np, scipy.linalg as sla, sys
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import bench_workloads as W, oracle as O
par = W.make_pendulum(num_points=256, M=500, shared_hypers=False)
cpu = W.build_oracle(par)
st = cpu.discretization.all_points
u = cpu.policy(st)
mean, err = cpu.dynamics(st, u)   # err = beta*sigma per output
vx = cpu.lyapunov_function(st); vm = cpu.lyapunov_function(mean)
lv = cpu.lipschitz_lyapunov(mean)
thr = cpu.threshold(st)
bound = np.sum(lv*err,axis=1,keepdims=True)
dec0 = (vm-vx)
neg = (dec0+bound < thr).ravel()
print("negative frac", neg.mean(), "fail w/o uncertainty", (dec0>=thr).mean(), "fail due to uncertainty only", ((dec0<thr)&~neg.reshape(-1,1)).mean())
print("err stats", err.min(), np.median(err), err.max())
# partial variance bound with first r training points
z = np.hstack([st,u])
for r in (64,128,192,256,384):
    errs=[]
    for j,f in enumerate(cpu.dynamics.functions):
        gp=f.gaussian_process
        X=np.asarray(gp.X); kern=gp.kern
        L=np.asarray(gp.cholesky)
        Kx = gp._scale**2*kern.K(X[:r], z)
        a = sla.solve_triangular(L[:r,:r], Kx, lower=True)
        var = (gp._scale**2*kern.Kdiag(z) - np.sum(a*a,axis=0))/gp._scale**2
        errs.append(f.beta*np.sqrt(np.maximum(var,0)))
    eu=np.stack(errs,axis=1)
    bu=np.sum(lv*eu,axis=1,keepdims=True)
    decT = (dec0+bu < thr).ravel()
    decF = (dec0>=thr).ravel()
    und = ~(decT|decF)
    # tile-level: 64 consecutive
    tile_und = und.reshape(-1,64).any(axis=1)
    print(r, "err ratio median", np.median(eu/err), "undecided pts", und.mean(), "tiles with any undecided", tile_und.mean())
