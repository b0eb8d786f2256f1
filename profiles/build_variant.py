"""
A/B builds of the library: python profiles/build_variant.py <name> [-DFLAG ...]
  -> peakondrake_b200/libpeakon_b200_<name>.so with only the M=8 step kernels (+ M=10 when PK_WITH_M10 is in the
     environment), compiled with the extra flags.  Use with PK_LIB_PATH=<that file> (peakondrake_b200/_lib.py).
"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
name, extra = sys.argv[1], sys.argv[2:]
d = os.path.join(ROOT, 'peakondrake_b200', 'csrc')
bd = os.path.join(ROOT, 'build', 'var_' + name)
os.makedirs(bd, exist_ok=True)
flags = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-Xcompiler', '-fPIC',
         '-diag-suppress', '128', '-DPK_PROFILE_SINGLE_M'] + extra
jobs = [('pk_lib.cu', 'pk_lib.o', []), ('pk_upwind_lib.cu', 'pk_upwind_lib.o', []), ('pk_aux_lib.cu', 'pk_aux_lib.o', []),
        ('pk_kernels_m.cu', 'pk_kernels_m8.o', ['-DPK_M=8'])]
if os.environ.get('PK_WITH_M10'):      # the wide-chunk cluster kernels of BASELINE config 5
    jobs.append(('pk_kernels_m.cu', 'pk_kernels_m10.o', ['-DPK_M=10']))
    flags.append('-DPK_VARIANT_WITH_M10')
procs = [subprocess.Popen(['/usr/local/cuda/bin/nvcc'] + flags + x + ['-c', '-o', os.path.join(bd, o), os.path.join(d, src)])
         for src, o, x in jobs]
assert all(p.wait() == 0 for p in procs)
out = os.path.join(ROOT, 'peakondrake_b200', 'libpeakon_b200_%s.so' % name)
subprocess.check_call(['/usr/local/cuda/bin/nvcc', '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', out] +
                      [os.path.join(bd, o) for _, o, _ in jobs])
print(out)
