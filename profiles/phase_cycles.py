"""
Per-phase cycle breakdown of pk_step_kernel (clock64 buckets of thread 0 of every CTA).
Builds a separate profiling library (-DPK_PROFILE), runs the bench workload once and prints the
share of each phase.  Run on the GPU box:  python profiles/phase_cycles.py
"""
import ctypes, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lib_path = os.path.join(ROOT, 'peakondrake_b200', 'libpeakon_b200_prof.so')
if not os.path.exists(lib_path) or os.environ.get('REBUILD'):
  subprocess.check_call(['/usr/local/cuda/bin/nvcc', '-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a',
                       '-lineinfo', '-Xcompiler', '-fPIC', '-shared', '-diag-suppress', '128', '-DPK_PROFILE',
                       '-DPK_M=8', '-DPK_PROFILE_SINGLE_M', '-o', lib_path,
                       os.path.join(ROOT, 'peakondrake_b200', 'csrc', 'pk_lib.cu'),
                       os.path.join(ROOT, 'peakondrake_b200', 'csrc', 'pk_upwind_lib.cu'),
                       os.path.join(ROOT, 'peakondrake_b200', 'csrc', 'pk_aux_lib.cu'),
                       os.path.join(ROOT, 'peakondrake_b200', 'csrc', 'pk_kernels_m.cu')])
from peakondrake_b200 import _lib
_lib.LIB_PATH = lib_path
import torch
from peakondrake_b200 import EnsembleIntegrator, build_initial_u
ne, R, K = int(os.environ.get('NE', 1000)), int(os.environ.get('R', 1024)), int(os.environ.get('K', 50))
fam = os.environ.get('FAM', 'sines'); nxi = 4 if fam == 'sines' else 1
eng = EnsembleIntegrator(ne, 40.0, float(os.environ.get('DT', 1e-3)), R, num_Xis=nxi, Xi_family=fam, sigma=0.2, seed=1)
eng.set_state(build_initial_u(ne, 40.0, 'proper_peak'))
DUMP = int(os.environ.get('DUMP', 0))
adv = (lambda: eng.step_dump(K, DUMP, suite=True)) if DUMP else (lambda: eng.step(K))
adv(); torch.cuda.synchronize()
eng.lib.pk_profile_enable.argtypes = [ctypes.c_void_p]; eng.lib.pk_profile_read.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
assert eng.lib.pk_profile_enable(eng.ctx) == 0
adv(); ms = eng.last_kernel_ms()
buf = np.zeros((R, 16), dtype=np.int64)
assert eng.lib.pk_profile_read(eng.ctx, buf.ctypes.data_as(ctypes.c_void_p)) == 0
names = ['dW draw', 'noise fill', 'residual-only check', 'assemble_begin', 'assemble_finish(+LU,tips)',
         'allreduce+test', 'reduce build/publish', 'PCR steps', 'PCR final', 'interior solve',
         'helm causal+scan', 'helm anticausal+scan+update', 'accept',
         'cluster: segment-local PCR', 'cluster: separator final + barrier', 'fused dump']
tot = buf.sum(1).mean()
print('kernel %.3f ms; mean cycles per CTA per step %.0f (its/step %.2f)' % (ms, tot / K, eng.its_total.double().mean().item() / (2 * K)))
for i, n in enumerate(names):
    print('%-30s %8.0f cyc/step  %5.1f%%' % (n, buf[:, i].mean() / K, 100 * buf[:, i].mean() / tot))
