"""
GPU parity tests: the CUDA path, called through the C ABI (ctypes -> libpeakon_b200.so), against
the NumPy/SciPy oracle on the same inputs and the same Brownian increments.

Tolerance (BASELINE.json north_star): relative L2 error of u per realisation <= 1e-10 in fp64;
Newton iteration counts and integer indices identical.
"""
import numpy as np
import pytest

from oracle import ch_oracle as O

pytestmark = pytest.mark.gpu

TOL = 1e-10


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def run_case(ne, R, nsteps, ic='proper_peak', fam='sines', nxi=4, sigma=0.2, scheme='hydrodynamic',
             Ld=40.0, dt=1e-3, peak_width=1.0, chunks=None):
    import torch
    from peakondrake_b200 import EnsembleIntegrator, build_initial_u, draw_dW
    u0 = build_initial_u(ne, Ld, ic, peak_width=peak_width)
    u0_or = O.initial_u(O.Mesh1D(ne, Ld), ic, peak_width=peak_width)
    assert rel(u0, u0_or) < 1e-12
    eng = EnsembleIntegrator(ne, Ld, dt, R, scheme=scheme, num_Xis=nxi, Xi_family=fam, sigma=sigma)
    eng.set_state(u0)
    dW = None
    if nxi > 0:
        dW = np.stack([draw_dW(r, nsteps, nxi, dt) for r in range(R)], axis=1)
    chunks = chunks or [nsteps]
    assert sum(chunks) == nsteps
    n0 = 0
    for c in chunks:           # several launches must equal one launch
        eng.step(c, dW=None if dW is None else dW[n0:n0 + c])
        n0 += c
    torch.cuda.synchronize()
    tabs = O.noise_tables(O.Mesh1D(ne, Ld), fam, nxi, sigma, scheme=scheme) if nxi else None
    u, u1, dF, dG = [t.cpu().numpy() for t in (eng.u, eng.u1, eng.dF, eng.dG)]
    its = eng.its_total.cpu().numpy()
    fs = eng.fail_step.cpu().numpy()
    worst = 0.0
    for r in range(R):
        ref = O.run(ne, Ld, dt, nsteps, u0, None if dW is None else dW[:, r, :], tabs, scheme=scheme)
        assert fs[r] == -1 and ref['fail_step'] == -1
        assert its[r] == ref['its'].sum(), (r, its[r], ref['its'].sum())
        worst = max(worst, rel(u[r], ref['u']))
        assert rel(u[r], ref['u']) <= TOL
        assert rel(u1[r], ref['w1'][:ne]) <= TOL
        assert rel(dF[r], ref['w1'][ne:2 * ne]) <= 1e-8
        if np.linalg.norm(ref['w1'][2 * ne:]) > 0:
            assert rel(dG[r], ref['w1'][2 * ne:]) <= 1e-8
        else:
            assert not dG[r].any()
    eng.close()
    return worst


def test_c1_deterministic_single_peak():
    # BASELINE config 1 (shortened): CG1 ne=1000 L=40 dt=1e-3, one_peak, sigma=0
    assert run_case(1000, 2, 20, ic='one_peak', nxi=1, sigma=0.0) <= TOL


def test_c2_stochastic_pathwise():
    # BASELINE config 2 (shortened): 4 sines, sigma=0.2, proper_peak, shared dW per seed
    assert run_case(1000, 8, 20, chunks=[7, 13]) <= TOL


@pytest.mark.parametrize('ne', [8, 20, 37, 200, 1001, 1500, 2000])
def test_mesh_sizes(ne):
    assert run_case(ne, 3, 6, nxi=2) <= TOL


def test_constant_noise_family():
    # Xi_family 'constant': dXi_x = dXi_xx = 0 and dG stays identically zero
    assert run_case(1000, 4, 10, ic='new_peak', fam='constant', nxi=1, sigma=0.2, dt=5e-4, peak_width=1 / 6) <= TOL


def test_large_single_cta_mesh_scalar_noise():
    # ne=4000 fits one CTA only with scalar noise (3 resident fields): T=512, M=8
    assert run_case(4000, 2, 4, ic='periodic_peakon', fam='constant', nxi=1, sigma=0.5) <= TOL


def test_unsupported_mesh_raises_cleanly():
    from peakondrake_b200 import EnsembleIntegrator, _lib
    with pytest.raises(_lib.PkError):
        EnsembleIntegrator(200000, 40.0, 1e-3, 2, num_Xis=4, sigma=0.1)    # beyond the largest cluster layout
    with pytest.raises(_lib.PkError):
        EnsembleIntegrator(6, 40.0, 1e-3, 2)                               # fewer than 8 cells
    with pytest.raises(ValueError):
        EnsembleIntegrator(100, 40.0, 1e-3, 2, scheme='conforming')         # equations.py:147-148
    with pytest.raises(NotImplementedError):
        EnsembleIntegrator(100, 40.0, 1e-3, 2, alphasq=0.0, gamma=1.0)      # equations.py:160-161


def test_no_gradient_scheme():
    assert run_case(400, 3, 8, scheme='no_gradient', nxi=3) <= TOL


@pytest.mark.parametrize('ne,fam,nxi,sigma', [(400, 'sines', 3, 0.2), (1000, 'sines', 4, 0.2), (300, 'constant', 1, 1.0),
                                             (64, 'sines', 0, 0.0), (2000, 'sines', 2, 0.1), (6000, 'sines', 2, 0.1),
                                             (20000, 'constant', 1, 0.5)])       # the last two: cluster kernels
def test_test_scheme(ne, fam, nxi, sigma):
    """Scheme 'test' / 'midpoint' (equations.py:132-145) against the oracle: same Newton counts, u within 1e-10."""
    import torch
    from peakondrake_b200 import EnsembleIntegrator, build_initial_u, draw_dW
    Ld, dt, nsteps, R = 40.0, 1e-3, 9, 3
    u0 = build_initial_u(ne, Ld, 'proper_peak')
    eng = EnsembleIntegrator(ne, Ld, dt, R, scheme='test', num_Xis=nxi, Xi_family=fam, sigma=sigma)
    eng.set_state(u0)
    dW = np.stack([draw_dW(r, nsteps, nxi, dt) for r in range(R)], axis=1) if nxi else None
    eng.step(4, dW=None if dW is None else dW[:4])
    eng.step(5, dW=None if dW is None else dW[4:])
    torch.cuda.synchronize()
    tabs = O.noise_tables(O.Mesh1D(ne, Ld), fam, nxi, sigma, scheme='test') if nxi else None
    u, u1 = eng.u.cpu().numpy(), eng.u1.cpu().numpy()
    for r in range(R):
        ref = O.run(ne, Ld, dt, nsteps, u0, None if dW is None else dW[:, r, :], tabs, scheme='test')
        assert eng.fail_step[r].item() == -1 and eng.its_total[r].item() == ref['its'].sum()
        assert rel(u[r], ref['u']) <= TOL and rel(u1[r], ref['w1'][:ne]) <= TOL
    assert not eng.dF.any().item()
    eng.close()


def test_no_noise_modes():
    assert run_case(300, 2, 8, nxi=0) <= TOL


def test_periodic_peakon_large_sigma():
    assert run_case(2000, 3, 6, ic='periodic_peakon', fam='constant', nxi=1, sigma=1.0, dt=1e-3) <= TOL


def test_philox_export_roundtrip_and_sharding():
    """In-kernel Philox increments: exported dW fed to the oracle reproduces the path, and the
    numbers depend only on the global realisation id (sharding-invariant)."""
    import torch
    from peakondrake_b200 import EnsembleIntegrator, build_initial_u
    ne, Ld, dt, R, nsteps, nxi, sigma = 200, 40.0, 1e-3, 6, 8, 3, 0.3
    u0 = build_initial_u(ne, Ld, 'proper_peak')
    eng = EnsembleIntegrator(ne, Ld, dt, R, num_Xis=nxi, sigma=sigma, seed=1234)
    eng.set_state(u0)
    dW = eng.step(nsteps, export_dW=True).cpu().numpy()
    u = eng.u.cpu().numpy()
    assert abs(dW.std() / np.sqrt(dt) - 1.0) < 0.25 and abs(dW.mean()) < 0.02
    tabs = O.noise_tables(O.Mesh1D(ne, Ld), 'sines', nxi, sigma)
    for r in range(R):
        ref = O.run(ne, Ld, dt, nsteps, u0, dW[:, r, :], tabs)
        assert rel(u[r], ref['u']) <= TOL
    # second shard holding realisations 3..5 only
    eng2 = EnsembleIntegrator(ne, Ld, dt, 3, num_Xis=nxi, sigma=sigma, seed=1234, rid_offset=3)
    eng2.set_state(u0)
    dW2 = eng2.step(nsteps, export_dW=True).cpu().numpy()
    assert np.array_equal(dW2, dW[:, 3:, :])
    assert np.array_equal(eng2.u.cpu().numpy(), u[3:])
    # nXi_updates: sum of sub-draws has the right variance
    eng3 = EnsembleIntegrator(ne, Ld, dt, 64, num_Xis=1, Xi_family='constant', sigma=0.1, seed=7, nXi_updates=16)
    eng3.set_state(u0)
    d3 = eng3.step(16, export_dW=True).cpu().numpy()
    assert abs(d3.std() / np.sqrt(dt) - 1.0) < 0.15
    torch.cuda.synchronize()


def test_host_buffer_entry_point_equals_device_path():
    from peakondrake_b200 import EnsembleIntegrator, build_initial_u, draw_dW
    ne, Ld, dt, R, nsteps, nxi = 200, 40.0, 1e-3, 5, 6, 2
    u0 = build_initial_u(ne, Ld, 'one_peak')
    dW = np.stack([draw_dW(r, nsteps, nxi, dt) for r in range(R)], axis=1)
    eng = EnsembleIntegrator(ne, Ld, dt, R, num_Xis=nxi, sigma=0.2)
    eng.set_state(u0)
    eng.step(nsteps, dW=dW)
    u_dev = eng.u.cpu().numpy()
    its_dev = eng.its_total.cpu().numpy().copy()
    eng.set_state(u0)
    u = np.tile(u0, (R, 1)); u1 = np.zeros_like(u); dF = np.zeros_like(u); dG = np.zeros_like(u)
    fs = np.full(R, -1, np.int32); il = np.zeros(R, np.int32); it = np.zeros(R, np.int32)
    eng.step_host(nsteps, u, u1, dF, dG, np.ascontiguousarray(dW), fs, il, it)
    assert np.array_equal(u, u_dev)
    assert (fs == -1).all() and (it == its_dev).all()


def test_host_entry_point_chunk_pipeline_and_resident_warm_start():
    """R >= 256 goes through the chunked H2D / kernel / D2H pipeline; u1 = dF = dG = None keeps the Newton unknown
    in the context (the reference's Equations object owns it, equations.py:71,89).  Two calls of 3 steps must equal
    the device path's 6 steps bit for bit, in both modes, with Philox streams independent of the chunking."""
    from peakondrake_b200 import EnsembleIntegrator, build_initial_u
    ne, Ld, dt, R, nxi = 150, 40.0, 1e-3, 300, 3
    u0 = build_initial_u(ne, Ld, 'proper_peak')
    eng = EnsembleIntegrator(ne, Ld, dt, R, num_Xis=nxi, sigma=0.2, seed=11)
    eng.set_state(u0)
    eng.step(6)
    u_dev, w_dev = eng.u.cpu().numpy(), [t.cpu().numpy() for t in (eng.u1, eng.dF, eng.dG)]
    its_dev = eng.its_total.cpu().numpy().copy()
    for resident in (False, True):
        eng.lib.pk_set_step_counter(eng.ctx, 0)
        eng.reset_warm_start()
        u = np.tile(u0, (R, 1)); u1 = np.zeros_like(u); dF = np.zeros_like(u); dG = np.zeros_like(u)
        fs = np.full(R, -1, np.int32); il = np.zeros(R, np.int32); it = np.zeros(R, np.int32)
        for _ in range(2):
            if resident:
                eng.step_host(3, u, None, None, None, None, fs, il, it)
            else:
                eng.step_host(3, u, u1, dF, dG, None, fs, il, it)
        assert np.array_equal(u, u_dev) and (it == its_dev).all() and (fs == -1).all()
        if not resident:
            assert all(np.array_equal(a, b) for a, b in zip((u1, dF, dG), w_dev))
    eng.close()


def test_failure_latch():
    """A step SNES cannot solve latches the realisation (simulation.py:99-106): u frozen at the last
    accepted value, fail_step recorded, diagnostics NaN (outputting.py:107-108); healthy
    realisations of the same batch are unaffected."""
    import torch
    from peakondrake_b200 import EnsembleIntegrator, build_initial_u, _lib
    ne, Ld, dt, R, nsteps = 200, 40.0, 1e-3, 3, 4
    u0 = build_initial_u(ne, Ld, 'proper_peak')
    eng = EnsembleIntegrator(ne, Ld, dt, R, num_Xis=1, Xi_family='sines', sigma=1.0)
    eng.set_state(u0)
    dW = np.zeros((nsteps, R, 1))
    dW[:, 0, 0] = 0.01
    dW[:, 2, 0] = 0.01
    dW[2, 1, 0] = np.nan          # realisation 1 receives a non-finite increment at step 2
    eng.step(nsteps, dW=dW)
    vals, idx, _ = eng.diagnostics()
    torch.cuda.synchronize()
    fs = eng.fail_step.cpu().numpy()
    assert list(fs) == [-1, 2, -1]
    tabs = O.noise_tables(O.Mesh1D(ne, Ld), 'sines', 1, 1.0)
    ref1 = O.run(ne, Ld, dt, 2, u0, dW[:2, 1, :], tabs)     # state after the two good steps
    assert rel(eng.u[1].cpu().numpy(), ref1['u']) <= TOL
    ref0 = O.run(ne, Ld, dt, nsteps, u0, dW[:, 0, :], tabs)
    assert rel(eng.u[0].cpu().numpy(), ref0['u']) <= TOL
    v = vals.cpu().numpy()
    assert np.isnan(v[1]).all() and np.isfinite(v[0, _lib.D_ENERGY]) and np.isfinite(v[2, _lib.D_ENERGY])


def test_diagnostics_against_oracle():
    import torch
    from peakondrake_b200 import EnsembleIntegrator, build_initial_u, draw_dW, _lib
    ne, Ld, dt, R, nsteps, nxi = 500, 40.0, 1e-3, 4, 10, 4
    u0 = build_initial_u(ne, Ld, 'proper_peak')
    dW = np.stack([draw_dW(r, nsteps, nxi, dt) for r in range(R)], axis=1)
    eng = EnsembleIntegrator(ne, Ld, dt, R, num_Xis=nxi, sigma=0.2)
    eng.set_state(u0)
    eng.step(nsteps, dW=dW)
    pq = np.array([[1.0, 10.0 + 0.1 * r] for r in range(R)])
    vals, idx, du = eng.diagnostics(pq=pq, periodic=True, want_m=True, want_du=True)
    torch.cuda.synchronize()
    vals, idx, du, u = vals.cpu().numpy(), idx.cpu().numpy(), du.cpu().numpy(), eng.u.cpu().numpy()
    orc = O.CHOracle(ne, Ld, dt)
    mesh = orc.mesh
    names = ['energy', 'l2_u', 'h1_u', 'mass_u', 'max_du', 'min_du', 'max_jump_local', 'p_pde', 'l2_m',
             'h1_m', 'mass_m', 'm_max', 'E_2', 'u_error_with_sde']
    for r in range(R):
        m = O.m_field(orc, u[r])
        usde = O.u_sde_field(mesh, pq[r, 0], pq[r, 1], 1.0, True)
        d = O.scalar_diagnostics(orc, u[r], names, m=m, u_sde=usde)
        v = vals[r]
        close = lambda a, b: abs(a - b) <= 1e-11 * max(1.0, abs(b))
        assert close(v[_lib.D_ENERGY], d['energy'])
        assert close(np.sqrt(v[_lib.D_L2U2]), d['l2_u'])
        assert close(np.sqrt(v[_lib.D_L2U2] + v[_lib.D_H1SEMI2]), d['h1_u'])
        assert close(v[_lib.D_MASS_U], d['mass_u'])
        assert v[_lib.D_MAX_U] == u[r].max()
        assert close(np.sqrt(v[_lib.D_L2M2]), d['l2_m'])
        assert close(np.sqrt(v[_lib.D_L2M2] + v[_lib.D_H1SEMI_M2]), d['h1_m'])
        assert close(v[_lib.D_MASS_M], d['mass_m'])
        assert close(v[_lib.D_M_MAX], d['m_max'])
        assert close(0.5 * v[_lib.D_UMU], d['E_2'])
        assert close(np.sqrt(v[_lib.D_ERR_SDE2]), d['u_error_with_sde'])
        # du on the device is (u[i+1]-u[i])*(1/h); indices are exact argmax/argmin of that field
        assert np.allclose(du[r], O.du_field(mesh, u[r]), rtol=1e-14, atol=0)
        assert v[_lib.D_MAX_DU] == du[r].max() and v[_lib.D_MIN_DU] == du[r].min()
        assert idx[r, _lib.I_ARGMAX_DU] == np.argmax(du[r])
        assert idx[r, _lib.I_ARGMIN_DU] == np.argmin(du[r])
        assert idx[r, _lib.I_ARGMAX_U] == np.argmax(u[r])
        jmp = np.abs(du[r] - np.roll(du[r], 1))
        assert v[_lib.D_MAX_JUMP_LOCAL] == jmp.max() and idx[r, _lib.I_ARGMAX_JUMP] == np.argmax(jmp)


def test_argmax_ties_take_lowest_index():
    import torch
    from peakondrake_b200 import EnsembleIntegrator, _lib
    ne = 64
    eng = EnsembleIntegrator(ne, 40.0, 1e-3, 2)
    u = np.zeros((2, ne))
    u[0, [5, 40, 41]] = 1.0          # three equal maxima
    u[1, [63, 0]] = 2.0              # maxima at both ends of the array
    eng.u.copy_(torch.from_numpy(u))
    vals, idx, du = eng.diagnostics(want_du=True)
    torch.cuda.synchronize()
    idx = idx.cpu().numpy(); du = du.cpu().numpy()
    assert idx[0, _lib.I_ARGMAX_U] == 5 and idx[1, _lib.I_ARGMAX_U] == 0
    for r in range(2):
        assert idx[r, _lib.I_ARGMAX_DU] == np.argmax(du[r]) and idx[r, _lib.I_ARGMIN_DU] == np.argmin(du[r])


def test_convergence_certificate_is_bit_neutral(monkeypatch):
    """PK_NO_CERTIFICATE=1 evaluates the last residual of every Newton solve like the reference does; the default
    decides the same test from |delta|^2.  Same iterates => bit-identical states and counters."""
    import torch
    from peakondrake_b200 import EnsembleIntegrator, build_initial_u
    out = []
    for off in ('0', '1'):
        monkeypatch.setenv('PK_NO_CERTIFICATE', off)
        eng = EnsembleIntegrator(1000, 40.0, 1e-3, 64, num_Xis=4, Xi_family='sines', sigma=0.2, seed=5)
        eng.set_state(build_initial_u(1000, 40.0, 'proper_peak'))
        eng.step(40); eng.step(23)
        torch.cuda.synchronize()
        out.append([t.cpu().numpy().copy() for t in (eng.u, eng.u1, eng.dF, eng.dG, eng.its_total, eng.fail_step)])
        eng.close()
    for a, b in zip(*out):
        assert np.array_equal(a, b)
    assert (out[0][5] == -1).all() and out[0][4].min() >= 63        # nobody failed; at least one iteration per step


@pytest.mark.parametrize('ne,Ld', [(1000, 40.0), (16384, 40.0), (777, 13.7), (2000, 1.0)])
def test_diag_kernel_division_free_slopes_are_ieee_quotients(ne, Ld):
    """pk_diag_kernel forms du = (u[i+1] - u[i]) / h without a division (q = x RN(1/h), exact remainder by FMA, one fused
    correction).  With the correctly rounded reciprocal that IS the correctly rounded quotient: compared bit for bit with
    NumPy's IEEE division on smooth, random, tiny, huge and signed-zero differences, together with the extrema and
    arg-indices that are built from it (outputting.py:128-142)."""
    import torch
    from peakondrake_b200 import EnsembleIntegrator, _lib
    rng = np.random.default_rng(ne)
    R = 6
    u = np.empty((R, ne))
    x = np.arange(ne) * (Ld / ne)
    u[0] = np.exp(-np.abs(x - Ld / 2))
    u[1] = rng.standard_normal(ne)
    u[2] = rng.standard_normal(ne) * 1e-150
    u[3] = rng.standard_normal(ne) * 1e150
    u[4] = np.round(rng.standard_normal(ne) * 4) / 4          # many exact ties and zero differences
    u[5] = np.ldexp(rng.integers(1, 2 ** 53, ne).astype(np.float64), rng.integers(-60, 60, ne))
    eng = EnsembleIntegrator(ne, Ld, 1e-3, R)
    eng.u.copy_(torch.from_numpy(u))
    vals, idx, du = eng.diagnostics(want_du=True)
    torch.cuda.synchronize()
    vals, idx, du = vals.cpu().numpy(), idx.cpu().numpy(), du.cpu().numpy()
    h = Ld / ne
    ref = (np.roll(u, -1, axis=1) - u) / h
    assert np.array_equal(du.view(np.int64), ref.view(np.int64))
    jump = np.abs(ref - np.roll(ref, 1, axis=1))
    for r in range(R):
        assert vals[r, _lib.D_MAX_DU] == ref[r].max() and vals[r, _lib.D_MIN_DU] == ref[r].min()
        assert idx[r, _lib.I_ARGMAX_DU] == np.argmax(ref[r]) and idx[r, _lib.I_ARGMIN_DU] == np.argmin(ref[r])
        assert vals[r, _lib.D_MAX_JUMP_LOCAL] == jump[r].max() and idx[r, _lib.I_ARGMAX_JUMP] == np.argmax(jump[r])
    eng.close()


# ---- thread-block-cluster variant (one realisation per cluster of CTAs, DSMEM) -------------------
@pytest.mark.parametrize('ne,C', [(1000, 2), (1000, 4), (2000, 8), (500, 2)])
def test_cluster_variant_forced_on_small_meshes(ne, C, monkeypatch):
    monkeypatch.setenv('PK_FORCE_CLUSTER', str(C))
    assert run_case(ne, 3, 6, nxi=2) <= TOL
    assert run_case(ne, 2, 4, ic='periodic_peakon', fam='constant', nxi=1, sigma=0.5) <= TOL


def test_c4_fine_mesh_16384():
    # BASELINE config 4 (shortened): steep front, ne=16384, constant noise, dt=5e-4
    from peakondrake_b200 import EnsembleIntegrator
    eng = EnsembleIntegrator(16384, 40.0, 5e-4, 2, num_Xis=1, Xi_family='constant', sigma=0.2)
    assert eng.layout['ctas_per_realisation'] == 8 and eng.layout['threads'] == 2048
    eng.close()
    assert run_case(16384, 2, 4, ic='new_peak', fam='constant', nxi=1, sigma=0.2, dt=5e-4, peak_width=1 / 6) <= TOL


def test_c5_fine_mesh_20000_large_noise():
    # BASELINE config 5 (shortened): periodic peakon, ne=20000, sigma=1, dt=1e-2 (3 Newton iterations/step)
    assert run_case(20000, 2, 3, ic='periodic_peakon', fam='constant', nxi=1, sigma=1.0, dt=1e-2) <= TOL


@pytest.mark.parametrize('ne,force,ic,sigma,dt', [(16384, 0, 'new_peak', 0.2, 5e-4), (20000, 0, 'periodic_peakon', 1.0, 1e-3),
                                                 (2000, 4, 'proper_peak', 0.5, 1e-3), (1000, 2, 'periodic_peakon', 0.5, 1e-3)])
def test_two_clusters_per_sm_kernels_bit_identical_to_two_buffer_kernels(ne, force, ic, sigma, dt, monkeypatch):
    """Scalar-noise cluster meshes run the two-CTAs-per-SM instantiations by default: ONE PCR row buffer, the tips and the
    next separator through CTA-local copies (block barriers instead of cluster barriers), all-gathers pushed with st.async +
    mbarrier, cluster barrier released by one warp.  PK_CLUSTER_MINB=1 selects the older kernels (two row buffers, every
    exchange behind a cluster barrier, DSMEM pulls).  Same arithmetic in the same order: bit-identical states and counters,
    over several launches (the mbarrier phases carry on across the steps of a launch and restart with each launch)."""
    import torch
    from peakondrake_b200 import EnsembleIntegrator, build_initial_u
    if force:
        monkeypatch.setenv('PK_FORCE_CLUSTER', str(force))
    out = []
    for minb in ('0', '1'):
        monkeypatch.setenv('PK_CLUSTER_MINB', minb)
        eng = EnsembleIntegrator(ne, 40.0, dt, 5, num_Xis=1, Xi_family='constant', sigma=sigma, seed=11)
        eng.set_state(build_initial_u(ne, 40.0, ic, peak_width=1 / 6))
        eng.step(7); eng.step(1); eng.step(4)
        torch.cuda.synchronize()
        out.append([t.cpu().numpy().copy() for t in (eng.u, eng.u1, eng.dF, eng.its_total, eng.fail_step)])
        if minb == '0':
            assert eng.layout['smem_bytes'] < 110000          # the single-buffer layout was chosen
        eng.close()
    for a, b in zip(*out):
        assert np.array_equal(a, b)
    assert (out[0][4] == -1).all() and out[0][3].min() >= 12


def test_cluster_with_noise_tables_4096():
    assert run_case(4096, 2, 4, nxi=4) <= TOL


def test_wide_chunk_cluster_layouts(monkeypatch):
    # meshes whose standard layout ends at 5 rows per thread run with 9 / 10 rows on half the threads
    from peakondrake_b200 import EnsembleIntegrator
    for ne, rows in ((20000, 10), (17000, 9)):
        eng = EnsembleIntegrator(ne, 40.0, 1e-3, 1, num_Xis=1, Xi_family='constant', sigma=0.2)
        assert eng.layout['threads'] == 2048 and eng.layout['rows_per_thread'] == rows and eng.layout['ctas_per_realisation'] == 8
        eng.close()
    assert run_case(17000, 2, 3, ic='periodic_peakon', fam='constant', nxi=1, sigma=0.5) <= TOL
    assert run_case(20000, 1, 2, nxi=4) <= TOL                 # 10 rows per thread with noise tables (7 fields in smem)
    # the standard layout and the older Schur-solve variants stay selectable for A/B runs and agree with the oracle too
    monkeypatch.setenv('PK_NO_WIDE_CHUNKS', '1')
    assert run_case(20000, 1, 2, ic='periodic_peakon', fam='constant', nxi=1, sigma=1.0, dt=1e-2) <= TOL
    monkeypatch.delenv('PK_NO_WIDE_CHUNKS')
    for var in ('PK_CLUSTER_FLAT_PCR', 'PK_CLUSTER_LEADER_SEP'):
        monkeypatch.setenv(var, '1')
        assert run_case(16384, 1, 2, ic='new_peak', fam='constant', nxi=1, sigma=0.2, dt=5e-4, peak_width=1 / 6) <= TOL
        monkeypatch.delenv(var)


# ---- upwind / ssprk3 (equations.py:28-61,172-185) ---------------------------------------------------
def run_upwind_case(ne, R, nsteps, ic='one_peak', fam='sines', nxi=7, sigma=0.25, dt=1e-3, Ld=40.0):
    import torch
    from peakondrake_b200 import EnsembleIntegrator, build_initial_u, draw_dW, _lib
    u0 = build_initial_u(ne, Ld, ic)
    eng = EnsembleIntegrator(ne, Ld, dt, R, scheme='upwind', timestepping='ssprk3', num_Xis=nxi, Xi_family=fam, sigma=sigma)
    eng.set_state(u0)
    orc = O.UpwindOracle(ne, Ld, dt)
    m_init = eng.m[0].cpu().numpy().copy()
    # m = u - a2 u_xx amplifies rounding by ~4 a2/h^2: the two set-up solves agree to cond * eps
    assert rel(m_init, orc.initial_m(u0)) <= 1e-15 * max(1e3, 4.0 / (Ld / ne) ** 2)
    dW = np.stack([draw_dW(r, nsteps, nxi, dt) for r in range(R)], axis=1) if nxi else None
    eng.step(nsteps // 2, dW=None if dW is None else dW[:nsteps // 2])
    eng.step(nsteps - nsteps // 2, dW=None if dW is None else dW[nsteps // 2:])
    vals, idx, _ = eng.diagnostics(want_m=True)
    torch.cuda.synchronize()
    xi = O.noise_tables(O.Mesh1D(ne, Ld), fam, nxi, sigma, scheme='upwind')[0] if nxi else None
    u, m, vals = eng.u.cpu().numpy(), eng.m.cpu().numpy(), vals.cpu().numpy()
    assert (eng.fail_step.cpu().numpy() == -1).all()
    for r in range(R):
        ref = O.run_upwind(ne, Ld, dt, nsteps, u0, None if dW is None else dW[:, r, :], xi, m0=m_init)
        assert rel(u[r], ref['u']) <= TOL and rel(m[r], ref['m']) <= TOL
        d = O.dg1_norms(orc, ref['m'])
        assert abs(np.sqrt(vals[r, _lib.D_L2M2]) - d['l2_m']) <= 1e-10 * d['l2_m']
        assert abs(np.sqrt(vals[r, _lib.D_L2M2] + vals[r, _lib.D_H1SEMI_M2]) - d['h1_m']) <= 1e-10 * d['h1_m']
        assert abs(vals[r, _lib.D_MASS_M] - d['mass_m']) <= 1e-10 and abs(vals[r, _lib.D_M_MAX] - d['m_max']) <= 1e-10
    eng.close()


def test_upwind_ssprk3_reference_test_parameters():
    # the reference's only runnable test (tests/test_dumping.py): 20 cells, 10 steps, 7 sines, sigma 0.25
    run_upwind_case(20, 2, 10)


@pytest.mark.parametrize('ne', [8, 37, 200, 1000, 2000, 4000, 6000, 10000, 16384, 20000])
def test_upwind_mesh_sizes(ne):     # beyond 4096 cells: one realisation per thread-block cluster (pk_upwind_cluster_kernel)
    # the explicit scheme is only stable for |us|/h well below 1: scale the noise with the mesh width,
    # otherwise rounding differences are amplified step by step and path-wise comparison is meaningless
    run_upwind_case(ne, 3, 8, nxi=3, sigma=0.25 * min(1.0, 200.0 / ne))


def test_upwind_constant_and_no_noise():
    run_upwind_case(500, 2, 8, fam='constant', nxi=1, sigma=0.3, ic='proper_peak')
    run_upwind_case(500, 2, 8, nxi=0)


def test_upwind_failure_on_nonfinite_increment():
    import torch
    from peakondrake_b200 import EnsembleIntegrator, build_initial_u
    ne = 100
    u0 = build_initial_u(ne, 40.0, 'one_peak')
    eng = EnsembleIntegrator(ne, 40.0, 1e-3, 2, scheme='upwind', timestepping='ssprk3', num_Xis=1, sigma=0.5)
    eng.set_state(u0)
    dW = np.full((4, 2, 1), 0.01); dW[2, 1, 0] = np.inf
    eng.step(4, dW=dW)
    torch.cuda.synchronize()
    assert list(eng.fail_step.cpu().numpy()) == [-1, 2]
    xi = O.noise_tables(O.Mesh1D(ne, 40.0), 'sines', 1, 0.5, scheme='upwind')[0]
    ref = O.run_upwind(ne, 40.0, 1e-3, 2, u0, dW[:2, 1, :], xi)
    assert rel(eng.u[1].cpu().numpy(), ref['u']) <= TOL          # frozen at the last good step


def test_c2_full_size_properties():
    """BASELINE config 2 at its full size (1024 realisations, ne=1000, 4 sines) through size-independent properties:
    the work queue / step-block split is invisible (one launch of 120 steps == 50 + 70), Philox increments exported by
    one engine and re-imported by two half-size shards reproduce every path bit for bit, nothing fails, Newton takes
    2 iterations per step almost always, the conserved quantity of the deterministic flow drifts only with the noise
    amplitude, and a sample of realisations matches the oracle to 1e-10."""
    import torch
    from peakondrake_b200 import EnsembleIntegrator, build_initial_u
    ne, Ld, dt, R, nxi, sigma, nsteps = 1000, 40.0, 1e-3, 1024, 4, 0.2, 120
    u0 = build_initial_u(ne, Ld, 'proper_peak')
    eng = EnsembleIntegrator(ne, Ld, dt, R, num_Xis=nxi, sigma=sigma, seed=99)
    eng.set_state(u0)
    dW = eng.step(nsteps, export_dW=True)
    lay = eng.layout
    assert lay['ctas'] <= lay['resident_ctas'] and lay['step_blocks'] * lay['steps_per_block'] >= nsteps
    u = eng.u.clone()
    assert (eng.fail_step == -1).all().item()
    its = eng.its_total.double() / nsteps
    assert 1.9 <= its.mean().item() <= 2.1 and its.max().item() <= 3.0
    # two launches == one launch
    eng.set_state(u0); eng.lib.pk_set_step_counter(eng.ctx, 0)
    eng.step(50); eng.step(70)
    assert torch.equal(eng.u, u)
    # sharding: two engines of 512 realisations fed with the exported increments
    for k in range(2):
        e2 = EnsembleIntegrator(ne, Ld, dt, R // 2, num_Xis=nxi, sigma=sigma, seed=99, rid_offset=k * (R // 2))
        e2.set_state(u0)
        e2.step(nsteps, dW=dW[:, k * (R // 2):(k + 1) * (R // 2)].contiguous())
        assert torch.equal(e2.u, u[k * (R // 2):(k + 1) * (R // 2)])
        if k == 1:      # and drawn in-kernel: same numbers because the counter is the global realisation id
            e2.set_state(u0); e2.lib.pk_set_step_counter(e2.ctx, 0)
            e2.step(nsteps)
            assert torch.equal(e2.u, u[R // 2:])
        e2.close()
    # oracle on a sample
    tabs = O.noise_tables(O.Mesh1D(ne, Ld), 'sines', nxi, sigma)
    dWh, uh = dW.cpu().numpy(), u.cpu().numpy()
    for r in (0, 517, 1023):
        ref = O.run(ne, Ld, dt, nsteps, u0, dWh[:, r, :], tabs)
        assert rel(uh[r], ref['u']) <= TOL
    eng.close()


def test_c3_full_size_shard_statistics():
    """BASELINE config 3, one GPU's shard (8192 realisations, ne=2000, periodic peakon, 4 sines): the packed
    ensemble statistics are consistent with the per-realisation diagnostics (integer histogram counts sum to the
    number of healthy realisations, the per-x mean equals the mean of the states, the second moment bounds it)."""
    import torch
    from peakondrake_b200 import EnsembleIntegrator, build_initial_u, ensemble as ens
    ne, Ld, dt, R = 2000, 40.0, 1e-3, 8192
    eng = EnsembleIntegrator(ne, Ld, dt, R, num_Xis=4, sigma=0.2, seed=5)
    eng.set_state(build_initial_u(ne, Ld, 'periodic_peakon'))
    eng.step(20)
    st = ens.EnsembleStatistics(eng, nbins=64, lo=0.0, hi=Ld)
    res = st.update(False)
    torch.cuda.synchronize()
    healthy = int((eng.fail_step < 0).sum().item())
    assert healthy == R
    packed = st.packed.cpu().numpy()
    s1, s2, n = packed[:ne], packed[ne:2 * ne], packed[2 * ne]
    hq, hd = packed[2 * ne + 1:2 * ne + 65], packed[2 * ne + 65:2 * ne + 129]
    assert n == R and hq.sum() == R and hd.sum() == R and (hq == np.round(hq)).all()
    um = eng.u.double().mean(0).cpu().numpy()
    assert np.allclose(s1 / n, um, rtol=1e-12, atol=1e-14)
    assert (s2 / n + 1e-15 >= (s1 / n) ** 2).all()
    eng.close()


def test_c1_full_length_and_c2_long_path():
    """BASELINE config 1 at its full length (ne=1000, dt=1e-3, T=5: 5000 steps, deterministic single peakon) and one
    stochastic C2 path over 1000 steps: the error against the oracle must not accumulate beyond 1e-10 and the
    Newton iteration totals must be identical over the whole run."""
    import torch
    from peakondrake_b200 import EnsembleIntegrator, build_initial_u, draw_dW
    ne, Ld, dt = 1000, 40.0, 1e-3
    # C1: sigma = 0 with one 'sines' mode, ic 'one_peak'
    nsteps = 5000
    u0 = build_initial_u(ne, Ld, 'one_peak')
    eng = EnsembleIntegrator(ne, Ld, dt, 1, num_Xis=1, Xi_family='sines', sigma=0.0)
    eng.set_state(u0)
    dW = draw_dW(0, nsteps, 1, dt)[:, None, :]
    for k in range(0, nsteps, 1000):
        eng.step(1000, dW=dW[k:k + 1000])
    torch.cuda.synchronize()
    ref = O.run(ne, Ld, dt, nsteps, u0, dW[:, 0, :], O.noise_tables(O.Mesh1D(ne, Ld), 'sines', 1, 0.0))
    assert eng.fail_step[0].item() == -1 and ref['fail_step'] == -1
    assert eng.its_total[0].item() == ref['its'].sum()
    err = rel(eng.u[0].cpu().numpy(), ref['u'])
    assert err <= TOL, err
    eng.close()
    # C2: one stochastic path, 1000 steps, the reference's MT19937 stream for seed 7
    nsteps, nxi, sigma = 1000, 4, 0.2
    u0 = build_initial_u(ne, Ld, 'proper_peak')
    eng = EnsembleIntegrator(ne, Ld, dt, 2, num_Xis=nxi, Xi_family='sines', sigma=sigma)
    eng.set_state(u0)
    dW = np.stack([draw_dW(s, nsteps, nxi, dt) for s in (7, 8)], axis=1)
    eng.step(nsteps, dW=dW)
    torch.cuda.synchronize()
    ref = O.run(ne, Ld, dt, nsteps, u0, dW[:, 0, :], O.noise_tables(O.Mesh1D(ne, Ld), 'sines', nxi, sigma))
    assert eng.its_total[0].item() == ref['its'].sum()
    err = rel(eng.u[0].cpu().numpy(), ref['u'])
    assert err <= TOL, err
    eng.close()
