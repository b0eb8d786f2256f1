"""
Host side of the reference's `Outputting` (outputting.py): the NetCDF result file and the few
diagnostics that are post-processing of device results.

  - DataFile: the `data.nc` schema of experiment.py:99-156 written as NetCDF-3 (scipy.io.netcdf_file;
    netCDF4/HDF5 is not available in this image).
  - mu_diagnostic: restatement of find_mus (outputting.py:260-370) working on NumPy arrays of one
    realisation; sequential multi-peak logic, run on the host from the device's du field.
  - peakon_suite_from_indices: turns the integer results of the device scans (pk_peakon_suite) into the
    eight numbers of peakon_diagnostics (outputting.py:373-450) with the host's coordinate array, so
    coordinates are looked up, never recomputed (bit-exact locations).
"""
import numpy as np
from scipy.io import netcdf_file

PEAKON_DIAGNOSTICS = ['peakon_loc', 'peakon_min_du', 'peakon_max_du', 'peakon_min_du_loc',
                      'peakon_max_du_loc', 'peakon_max_u', 'peakon_mu', 'peakon_nu']

_NC_TYPES = {float: 'd', int: 'i', 'S1': 'c'}


class DataFile(object):
    """In-memory image of results/<code>/data[_dtI].nc, flushed to disk with save()."""

    def __init__(self, file_name):
        self.file_name = file_name
        self.dims = {}          # name -> size (None = unlimited)
        self.vars = {}          # name -> (nc type char, dims tuple)
        self.data = {}          # name -> ndarray (record variables grow along axis 0)

    def createDimension(self, name, size):
        self.dims[name] = size

    def createVariable(self, name, typ, dims):
        tc = _NC_TYPES.get(typ, 'd')
        if len(tuple(dims)) == 0:
            # scipy's NetCDF-3 writer corrupts record variables when a 0-d variable is present;
            # scalars are stored with a length-1 dimension instead (read as var[0], like the
            # reference's own `data_file['failed_time'][0]`, outputting.py:241)
            if 'scalar' not in self.dims:
                self.createDimension('scalar', 1)
            dims = ('scalar',)
        self.vars[name] = (tc, tuple(dims))
        shape = tuple(0 if self.dims[d] is None else self.dims[d] for d in dims)
        if tc == 'c':
            self.data[name] = np.full(shape, b' ', dtype='S1')
        elif tc == 'i':
            self.data[name] = np.zeros(shape, dtype=np.int32)
        else:
            self.data[name] = np.full(shape, np.nan)

    def _grow(self, name, n):
        a = self.data[name]
        if a.shape[0] < n:
            pad = np.full((n - a.shape[0],) + a.shape[1:], np.nan if a.dtype.kind == 'f' else 0, dtype=a.dtype)
            self.data[name] = np.concatenate([a, pad], axis=0)

    def put(self, name, index, value):
        """data[name][index] = value; index[0] may exceed the current length of an unlimited axis."""
        _, dims = self.vars[name]
        if dims and self.dims[dims[0]] is None:
            t = index[0] if isinstance(index, tuple) else index
            self._grow(name, int(t) + 1)
        self.data[name][index] = value

    def save(self):
        f = netcdf_file(self.file_name, 'w')
        nrec = max([self.data[v].shape[0] for v, (_, d) in self.vars.items() if d and self.dims[d[0]] is None] or [0])
        for name, size in self.dims.items():
            f.createDimension(name, size)
        handles = {}
        for name, (tc, dims) in self.vars.items():
            handles[name] = f.createVariable(name, tc, dims)
        for name, (tc, dims) in self.vars.items():
            a = self.data[name]
            if dims and self.dims[dims[0]] is None:
                self._grow(name, nrec)
                a = self.data[name]
                for t in range(nrec):
                    handles[name][t] = a[t]
            else:
                handles[name][:] = a
        f.close()


def read_data_file(file_name):
    """Load every variable of a NetCDF-3 result file into a dict of arrays."""
    f = netcdf_file(file_name, 'r', mmap=False)
    out = {k: np.array(v.data) for k, v in f.variables.items()}
    f.close()
    return out


def peakon_suite_from_indices(x_cells, max_u, vals, ids):
    """
    x_cells: DG0 coordinates (cell mid-points); max_u = max(u); vals = (min_du, max_du);
    ids = (argmax u, min_du_idx, max_du_idx) from pk_peakon_suite (-1 = not found).
    Mirrors the arithmetic of outputting.py:381-399,437-448.
    """
    n = len(x_cells)
    L = np.max(x_cells) + np.min(x_cells)
    pc = int(ids[0])
    xm, xp = x_cells[pc - 1], x_cells[pc]             # pc-1 == -1 wraps to the last cell, as in the reference
    peak_loc = 0.5 * (xm + xp - L) if xm > xp else 0.5 * (xm + xp)
    min_du, max_du = float(vals[0]), float(vals[1])
    x_min = x_cells[int(ids[1])] if ids[1] >= 0 else np.nan
    x_max = x_cells[int(ids[2])] if ids[2] >= 0 else np.nan
    mu = (x_min - x_max) if (x_max < x_min) else (L - x_min + x_max)
    nu = (max_du - min_du) / mu
    return {'peakon_loc': peak_loc, 'peakon_min_du': min_du, 'peakon_max_du': max_du,
            'peakon_min_du_loc': x_min, 'peakon_max_du_loc': x_max, 'peakon_max_u': float(max_u),
            'peakon_mu': mu, 'peakon_nu': nu}


def _cg1_at(u, Ld, x):
    """CG1 point evaluation on the periodic mesh (what Function.at does for a CG1 field)."""
    n = len(u)
    s = x / (Ld / n)
    i = int(np.floor(s)) % n
    th = s - np.floor(s)
    return (1 - th) * u[i] + th * u[(i + 1) % n]


def mu_diagnostic(u, du, x_cells, Ld):
    """
    The `mu` diagnostic: (old_mus, alt_mus) of find_mus (outputting.py:260-370) for one realisation.
    old: distance between the max and the min of du inside each peak's zone; alt: distance between
    the peak and the min of du.  Zones, the "four highest peaks" rule, the wrap-around cases and the
    half-mu fallback follow the reference statement by statement (including :316's comparison of a
    dof index with a peak count, App. B).
    """
    n = len(du)
    L = np.max(x_cells)
    wrap = lambda d: d if d < L / 2 else L - d
    mean2 = 2 * np.mean(u)
    prev = np.roll(du, 1)
    cand = np.nonzero((prev > 0) & (du < 0))[0]              # sign change + -> - between cells i-1 and i
    peaks = []                                               # (index, x_i, mid-point, height)
    for i in cand:
        mid = 0.5 * (x_cells[i - 1] + x_cells[i])
        height = _cg1_at(u, Ld, mid)
        if height > mean2:
            peaks.append((int(i), x_cells[i], mid, height))
    npk = len(peaks)
    old, alt = [], []
    if npk > 1:
        ranked = sorted((p[3] for p in peaks), reverse=True)
        cuts = []
        for k in range(npk):
            a, b = peaks[k - 1], peaks[k]
            if b[1] > a[1]:
                cuts.append(int(np.floor(0.5 * (a[0] + b[0]))))
            else:
                cuts.append(int(np.floor(0.5 * (a[0] + b[0] + n))) % n)
        for k in range(npk):
            if not (npk < 4 or peaks[k][3] > ranked[3]):
                continue
            lo, hi = cuts[k - 1], cuts[k]
            if hi > lo:
                imax = int(np.argmax(du[lo:hi])) + lo
                imin = int(np.argmin(du[lo:hi])) + lo
            elif hi == 0:
                imax = int(np.argmax(du[lo:])) + lo
                imin = int(np.argmin(du[lo:])) + lo
            elif lo == npk - 1:
                imax = int(np.argmax(du[0:lo])) + lo
                imin = int(np.argmin(du[0:lo])) + lo
            else:
                a1, a2 = int(np.argmax(du[0:hi])), int(np.argmax(du[lo:])) + lo
                b1, b2 = int(np.argmin(du[0:hi])), int(np.argmin(du[lo:])) + lo
                imax = a1 if du[a1] > du[a2] else a2
                imin = b1 if du[b1] < du[b2] else b2
            xmax, xmin = x_cells[imax], x_cells[imin]
            old.append(wrap(abs(xmin - xmax)))
            alt.append(wrap(abs(xmin - peaks[k][2])))
    else:
        xmax, xmin = x_cells[int(np.argmax(du))], x_cells[int(np.argmin(du))]
        old.append(wrap(abs(xmin - xmax)))
        if npk > 0:
            alt.append(wrap(abs(xmin - peaks[0][2])))
        else:
            print('No peakon detected, using half-mu instead')
            alt.append(wrap(abs(xmin - xmax)) / 2)
    return old, alt


# ---------------------------------------------------------------------------------------------------
# Field dumps: the reference writes `dirname/fields[_<dim><index>...]_output.pvd` through Firedrake's
# File (outputting.py:59-68,224-232), i.e. a ParaView collection of .vtu pieces, one per dump.  On a
# periodic interval Firedrake's output space is DG1 (every cell owns its two end points), which holds
# CG1, DG0 and DG1 fields alike; the same layout is written here as ASCII VTK XML.
# ---------------------------------------------------------------------------------------------------


def m_from_u(u, Ld, alphasq):
    """Diagnostic field m (diagnostic_equations.py:236-243): M m = (M + alphasq K) u on the periodic CG1 mesh."""
    from peakondrake_b200.stochastic_functions import circulant_tridiag_solve
    u = np.asarray(u, dtype=float)
    h = Ld / u.shape[-1]
    hb, ha = h / 6.0 - alphasq / h, 2.0 * h / 3.0 + 2.0 * alphasq / h
    b = hb * (np.roll(u, 1, -1) + np.roll(u, -1, -1)) + ha * u
    return circulant_tridiag_solve(h / 6.0, 2.0 * h / 3.0, b)


def u_sde_field(x, Ld, alphasq, p_true, q, periodic):
    """Nodal values of the peakon profile (diagnostic_equations.py:254-279) with the amplitude of :377; the
    third periodic branch keeps the reference's parenthesis as written (:270)."""
    a = np.sqrt(alphasq)
    x = np.asarray(x, dtype=float)
    E = np.exp(-Ld / a)
    p = p_true * 0.5 * (1 + E) / (1 - E)
    with np.errstate(over='ignore'):
        if periodic:
            b1 = p * ((np.exp(-(x - q + Ld) / a) + E * np.exp((x - q + Ld) / a)) / (1 - E))
            b2 = p * ((np.exp(-np.sqrt((q - x) ** 2 / alphasq)) + E * np.exp(np.sqrt((q - x) ** 2 / alphasq))) / (1 - E))
            b3 = p * ((np.exp(-(q + Ld - x) / a) + np.exp(-Ld / a * np.exp((q + Ld - x) / a))) / (1 - E))
        else:
            b1 = p * np.exp(-(x - q + Ld) / a)
            b2 = p * np.exp(-np.sqrt((q - x) ** 2 / alphasq))
            b3 = p * np.exp(-(q + Ld - x) / a)
    return np.where(x < q - Ld / 2, b1, np.where(x < q + Ld / 2, b2, b3))


def field_file_name(dirname, dim_names, indices):
    name = dirname + '/fields'
    for d, i in zip(dim_names, indices):
        name += '_' + str(d) + str(i)                                       # outputting.py:61-63
    return name + '_output.pvd'


def to_dg1(values, space):
    """[ne] CG1 / DG0 values or [ne, 2] DG1 values -> [ne, 2] (left end, right end) of every cell."""
    v = np.asarray(values, dtype=float)
    if space == 'DG1':
        return v.reshape(-1, 2)
    if space == 'DG0':
        return np.stack([v, v], axis=1)
    return np.stack([v, np.roll(v, -1)], axis=1)                           # CG1: cell i = [node i, node i+1]


class FieldFile(object):
    """`File(name.pvd)` + `.write(*fields, time=t)` of the reference (outputting.py:63,231)."""

    def __init__(self, pvd_name, ne, Ld):
        self.pvd_name, self.ne, self.Ld = pvd_name, int(ne), float(Ld)
        self.base = pvd_name[:-4]
        self.entries = []

    def write(self, fields, time):
        """fields: ordered mapping name -> (values, 'CG1' | 'DG0' | 'DG1')."""
        ne, h = self.ne, self.Ld / self.ne
        piece = '%s_%d.vtu' % (self.base, len(self.entries))
        xl = np.arange(ne) * h
        pts = np.stack([xl, xl + h], axis=1).ravel()
        with open(piece, 'w') as f:
            f.write('<?xml version="1.0" ?>\n<VTKFile type="UnstructuredGrid" version="0.1" byte_order="LittleEndian">\n')
            f.write('<UnstructuredGrid>\n<Piece NumberOfPoints="%d" NumberOfCells="%d">\n' % (2 * ne, ne))
            f.write('<Points>\n<DataArray type="Float64" NumberOfComponents="3" format="ascii">\n')
            f.write(' '.join('%.17g 0 0' % x for x in pts))
            f.write('\n</DataArray>\n</Points>\n<Cells>\n<DataArray type="Int32" Name="connectivity" format="ascii">\n')
            f.write(' '.join('%d %d' % (2 * i, 2 * i + 1) for i in range(ne)))
            f.write('\n</DataArray>\n<DataArray type="Int32" Name="offsets" format="ascii">\n')
            f.write(' '.join(str(2 * (i + 1)) for i in range(ne)))
            f.write('\n</DataArray>\n<DataArray type="UInt8" Name="types" format="ascii">\n')
            f.write(' '.join(['3'] * ne))                                   # VTK_LINE
            f.write('\n</DataArray>\n</Cells>\n<PointData>\n')
            for name, (vals, space) in fields.items():
                f.write('<DataArray type="Float64" Name="%s" NumberOfComponents="1" format="ascii">\n' % name)
                f.write(' '.join('%.17g' % v for v in to_dg1(vals, space).ravel()))
                f.write('\n</DataArray>\n')
            f.write('</PointData>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>\n')
        self.entries.append((float(time), piece.split('/')[-1]))
        with open(self.pvd_name, 'w') as f:
            f.write('<?xml version="1.0" ?>\n<VTKFile type="Collection" version="0.1" byte_order="LittleEndian">\n<Collection>\n')
            for t, name in self.entries:
                f.write('<DataSet timestep="%.17g" part="0" file="%s" />\n' % (t, name))
            f.write('</Collection>\n</VTKFile>\n')


def read_field_piece(vtu_name):
    """Test helper: {'points': [2 ne], name: [ne, 2]} of one .vtu piece written by FieldFile."""
    import re
    txt = open(vtu_name).read()
    res = {}
    m = re.search(r'<Points>\s*<DataArray[^>]*>\s*(.*?)\s*</DataArray>', txt, re.S)
    res['points'] = np.array(m.group(1).split(), dtype=float).reshape(-1, 3)[:, 0]
    pd = txt[txt.index('<PointData>'):]
    for m in re.finditer(r'<DataArray type="Float64" Name="([^"]+)"[^>]*>\s*(.*?)\s*</DataArray>', pd, re.S):
        res[m.group(1)] = np.array(m.group(2).split(), dtype=float).reshape(-1, 2)
    return res
