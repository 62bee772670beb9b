import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, cases, oracle, smash_b200
import test_gpu_structures as T
def viol(a,b,c=None):
    a,b=np.asarray(a,np.float64),np.asarray(b,np.float64)
    tol=1e-4+2e-3*np.abs(b); d=np.abs(a-b)
    ok=d<=tol
    if c is not None:
        c=np.asarray(c,np.float64); ok|=np.abs(a-c)<=1e-4+2e-3*np.abs(c)+np.abs(b-c)
    i=np.unravel_index(np.argmax(np.where(ok,0,d)),d.shape)
    return int((~ok).sum()), float(d.max()), i, float(a[i]), float(b[i]), (float(c[i]) if c is not None else None)
for st in ("vic-a",):
    for rnd in (False, True):
        a,b=T.pair(cases.cance, st, random=rnd, jobs_fun=("nse","kge")); c=T.exact(cases.cance, st, random=rnd)
        print(st, rnd, "qsim", viol(a.output.qsim,b.output.qsim,c.output.qsim), "cost", float(a.output.cost), float(b.output.cost), float(c.output.cost))
        for n in ("husl1","husl2","hlsl","hlr"):
            x,y,z=getattr(a.output.fstates,n),getattr(b.output.fstates,n),getattr(c.output.fstates,n)
            print("   ", n, float(np.abs(x-y).max()), float(np.abs(y-z).max()), float(np.abs(y).max()))
lib=smash_b200._lib.lib(); lib.smash_b200_set_option(b"math",0)
try:
    a,b=T.pair(cases.cance,"vic-a",random=True)
    print("math0 qsim", viol(a.output.qsim,b.output.qsim))
    for n in ("husl1","husl2","hlsl","hlr"):
        x,y=getattr(a.output.fstates,n),getattr(b.output.fstates,n); print("   ", n, float(np.abs(x-y).max()))
except Exception as e: print("math0 failed:", e)
lib.smash_b200_set_option(b"math",1)
def make():
    m=cases.cance(sparse=True,T=480); m.setup.save_qsim_domain=True; m.setup.save_net_prcp_domain=True; return m
for st in ("gr-c","vic-a"):
    try:
        a,b=T.pair(make,st); c=T.exact(make,st)
        print(st,"sparse qdom", viol(a.output.sparse_qsim_domain,b.output.sparse_qsim_domain,c.output.sparse_qsim_domain))
        print(st,"sparse netp", viol(a.output.sparse_net_prcp_domain,b.output.sparse_net_prcp_domain,c.output.sparse_net_prcp_domain))
    except Exception as e: print(st, "sparse failed:", repr(e))
# multiple run
m=cases.cance(T=480); m.setup.structure="gr-c"
names=("ci","cp","cst","lr")
ind=np.array([1+smash_b200.solver._derived_types.GPARAMETERS_NAME.index(n) for n in names],np.int32)
rng=np.random.default_rng(5)
smp=np.asfortranarray(np.stack([rng.uniform(0.5,5,6),rng.uniform(50,600,6),rng.uniform(100,2000,6),rng.uniform(1,30,6)]).astype(np.float32))
cost=np.zeros(6,np.float32); qsim=np.zeros((3,480,6),np.float32,order="F")
smash_b200.compute_multiple_run(m.setup,m.mesh,m.input_data,m.parameters,m.states,m.output,smp,ind,cost,qsim)
wc=np.zeros(6,np.float32); wq=np.zeros((3,480,6),np.float32,order="F")
oracle.compute_multiple_run(m.setup,m.mesh,m.input_data,m.parameters,m.states,m.output,smp,ind,wc,wq)
print("multi cost", cost, wc); print("multi q", viol(qsim,wq))
