import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from peakondrake_b200 import EnsembleIntegrator, pure_xi_tables
ne, R = 20000, 1024
t0=time.time(); eng = EnsembleIntegrator(ne, 40.0, 1e-6, R, num_Xis=1, Xi_family='constant', sigma=1.0, seed=5); print('engine', time.time()-t0)
t0=time.time(); tabs = [torch.from_numpy(np.ascontiguousarray(t)).to(eng.device) for t in pure_xi_tables(ne, 40.0, 'constant', 1, 1.0)]; print('tables', time.time()-t0)
mk = lambda v: torch.full((R,), v, dtype=torch.float64, device=eng.device)
E = float(np.exp(-40.0))
for n in (1000, 10000):
    p, q = mk(0.7), mk(11.3)
    torch.cuda.synchronize(); t0 = time.time()
    eng.peakon_sde_step(n, p, q, tabs, E, 'ito_euler', dW=None, step0=0)
    torch.cuda.synchronize(); print('philox', n, time.time() - t0)
    dW = torch.randn(n, R, 1, dtype=torch.float64, device=eng.device) * 1e-3
    p, q = mk(0.7), mk(11.3)
    torch.cuda.synchronize(); t0 = time.time()
    eng.peakon_sde_step(n, p, q, tabs, E, 'ito_euler', dW=dW)
    torch.cuda.synchronize(); print('imported', n, time.time() - t0)
