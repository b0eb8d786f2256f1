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
from collections import OrderedDict
import numpy as np
from scipy.io import netcdf_file

PEAKON_DIAGNOSTICS = ['peakon_loc', 'peakon_min_du', 'peakon_max_du', 'peakon_min_du_loc',
                      'peakon_max_du_loc', 'peakon_max_u', 'peakon_mu', 'peakon_nu']

_NC_TYPES = {float: 'd', int: 'i', 'S1': 'c'}
# NetCDF classic (CDF-1/2) type codes, big-endian numpy types, element sizes
_NC_CODE = {'d': (6, '>f8', 8), 'i': (4, '>i4', 4), 'c': (2, 'S1', 1)}


def _pad4(n):
    return (n + 3) & ~3


class DataFile(object):
    """
    results/<code>/data[_dtI].nc written as NetCDF-3 (64-bit-offset classic format, CDF-2) by a small native
    writer, so that the file GROWS like the reference's netCDF4 'a'-mode file does (outputting.py:32,198-220):

      * createDimension / createVariable build the schema (experiment.py:99-156); the header is laid out on the
        first write and is final from then on;
      * a dump writes ONE record (all record variables at that time index, one contiguous slab) at its place at
        the end of the file and updates the record count -- O(record) per dump, nothing is rewritten, and every
        dump already made survives a later crash or ConvergenceError;
      * fixed-size variables (x, deltax, swept parameter values, wallclock_time, failed_time) are written in
        place by sync().

    An in-memory mirror (`data[name]`, record variables grow along axis 0) keeps the put()/data[...] interface and
    lets later batches of an experiment fill other slices of an existing record (the record is then re-written).
    The result is readable by any NetCDF-3 reader (tests read it back with scipy.io.netcdf_file).
    """

    def __init__(self, file_name):
        self.file_name = file_name
        self.dims = OrderedDict()   # name -> size (None = unlimited)
        self.vars = OrderedDict()   # name -> (nc type char, dims tuple)
        self.data = {}              # name -> ndarray
        self._fh = None
        self._begin = {}            # name -> file offset (record variables: offset inside the first record)
        self._recsize = 0
        self._rec0 = 0
        self._nrec_disk = 0
        self._dirty = set()         # record indices changed since the last flush

    @classmethod
    def reopen(cls, file_name):
        """Open an existing result file for appending (the reference's Dataset(file_name, 'a'), outputting.py:32):
        the schema and contents are loaded; the first sync() lays the file out again."""
        f = netcdf_file(file_name, 'r', mmap=False)          # raises if the file does not exist
        self = cls(file_name)
        for name, size in f.dimensions.items():
            self.dims[name] = size
        nrec = 0
        rev = {'d': float, 'i': int, 'c': 'S1'}
        for name, v in f.variables.items():
            tc = v.typecode()
            self.createVariable(name, rev.get(tc, float), v.dimensions)
            a = np.array(v.data)
            if a.dtype.kind in 'fi':
                a = a.astype(a.dtype.newbyteorder('='))
            self.data[name] = a.reshape(v.shape) if v.shape else a
            if self._is_rec(name):
                nrec = max(nrec, a.shape[0])
        f.close()
        self._nrec = nrec
        self._dirty = set(range(nrec))
        return self

    # ---- schema ----
    def createDimension(self, name, size):
        if self._fh is not None:
            raise RuntimeError('DataFile: the schema is final once data has been written')
        self.dims[name] = size

    def createVariable(self, name, typ, dims):
        if self._fh is not None:
            raise RuntimeError('DataFile: the schema is final once data has been written')
        tc = _NC_TYPES.get(typ, 'd')
        dims = tuple(dims)
        self.vars[name] = (tc, dims)
        shape = tuple(0 if self.dims[d] is None else self.dims[d] for d in dims)
        if tc == 'c':
            self.data[name] = np.full(shape, b' ', dtype='S1')
        elif tc == 'i':
            self.data[name] = np.zeros(shape, dtype=np.int32)
        else:
            self.data[name] = np.full(shape, np.nan)

    def _is_rec(self, name):
        dims = self.vars[name][1]
        return bool(dims) and self.dims[dims[0]] is None

    def _grow(self, name, n):
        a = self.data[name]
        if a.shape[0] < n:
            cap = max(n, 2 * a.shape[0])            # amortised growth; `nrec` tracks the logical length
            fill = np.nan if a.dtype.kind == 'f' else (b' ' if a.dtype.kind == 'S' else 0)
            pad = np.full((cap - a.shape[0],) + a.shape[1:], fill, dtype=a.dtype)
            self.data[name] = np.concatenate([a, pad], axis=0)

    @property
    def nrec(self):
        return getattr(self, '_nrec', 0)

    def _touch(self, name, t):
        t = int(t)
        self._grow(name, t + 1)
        if t + 1 > self.nrec:
            self._nrec = t + 1
        self._dirty.add(t)

    # ---- data ----
    def put(self, name, index, value):
        """data[name][index] = value; index[0] may exceed the current length of an unlimited axis."""
        if self._is_rec(name):
            self._touch(name, index[0] if isinstance(index, tuple) else index)
        self.data[name][index] = value

    def put_block(self, name, index, values):
        """Vectorised put: `index` may contain slices (the realisation axis of a batch): ONE array assignment per
        diagnostic and dump instead of one call per realisation."""
        self.put(name, index, values)

    def view(self, name):
        """The logical contents of a variable (record variables cut to the number of records written)."""
        a = self.data[name]
        return a[:self.nrec] if self._is_rec(name) else a

    # ---- file layout ----
    def _name_bytes(self, s):
        b = s.encode()
        return np.array(len(b), '>i4').tobytes() + b + b'\0' * (_pad4(len(b)) - len(b))

    def _vsize(self, name):
        tc, dims = self.vars[name]
        n = 1
        for d in dims:
            if self.dims[d] is not None:
                n *= self.dims[d]
        return _pad4(n * _NC_CODE[tc][2])

    def _layout(self):
        names = list(self.vars)
        dim_ids = {d: i for i, d in enumerate(self.dims)}
        fixed = [v for v in names if not self._is_rec(v)]
        recs = [v for v in names if self._is_rec(v)]

        def header(begins, nrec):
            h = b'CDF\x02' + np.array(nrec, '>i4').tobytes()
            if self.dims:
                h += np.array([0x0A, len(self.dims)], '>i4').tobytes()
                for d, size in self.dims.items():
                    h += self._name_bytes(d) + np.array(0 if size is None else size, '>i4').tobytes()
            else:
                h += np.zeros(2, '>i4').tobytes()
            h += np.zeros(2, '>i4').tobytes()                       # no global attributes
            if names:
                h += np.array([0x0B, len(names)], '>i4').tobytes()
                for v in names:
                    tc, dims = self.vars[v]
                    h += self._name_bytes(v) + np.array([len(dims)] + [dim_ids[d] for d in dims], '>i4').tobytes()
                    h += np.zeros(2, '>i4').tobytes()               # no attributes
                    h += np.array([_NC_CODE[tc][0], self._vsize(v)], '>i4').tobytes()
                    h += np.array(begins.get(v, 0), '>i8').tobytes()
            else:
                h += np.zeros(2, '>i4').tobytes()
            return h

        hlen = len(header({}, 0))
        off, begins = hlen, {}
        for v in fixed:
            begins[v] = off
            off += self._vsize(v)
        self._rec0 = off
        rel = 0
        for v in recs:
            begins[v] = off + rel
            rel += self._vsize(v)
        self._recsize = rel
        self._begin, self._fixed, self._recs, self._header = begins, fixed, recs, header

    def _open(self):
        if self._fh is None:
            self._layout()
            self._fh = open(self.file_name, 'w+b')
            self._fh.write(self._header(self._begin, 0))
            self._write_fixed()

    def _write_fixed(self):
        fh = self._fh
        for v in self._fixed:
            tc = self.vars[v][0]
            raw = np.ascontiguousarray(self.data[v]).astype(_NC_CODE[tc][1]).tobytes()
            fh.seek(self._begin[v])
            fh.write(raw + b'\0' * (self._vsize(v) - len(raw)))

    def _record_bytes(self, t):
        parts = []
        for v in self._recs:
            tc = self.vars[v][0]
            a = self.data[v]
            self._grow(v, t + 1)
            raw = np.ascontiguousarray(self.data[v][t]).astype(_NC_CODE[tc][1]).tobytes()
            parts.append(raw + b'\0' * (self._vsize(v) - len(raw)))
        return b''.join(parts)

    def sync(self):
        """Write what changed since the last call: dirty records (appended or re-written in place), the record
        count and the fixed-size variables."""
        self._open()
        fh = self._fh
        for t in sorted(self._dirty):
            fh.seek(self._rec0 + t * self._recsize)
            fh.write(self._record_bytes(t))
        if self.nrec > self._nrec_disk:
            # records never touched (gaps) must exist on disk too
            for t in range(self._nrec_disk, self.nrec):
                if t not in self._dirty:
                    fh.seek(self._rec0 + t * self._recsize)
                    fh.write(self._record_bytes(t))
            self._nrec_disk = self.nrec
            fh.seek(4)
            fh.write(np.array(self.nrec, '>i4').tobytes())
        self._dirty.clear()
        self._write_fixed()
        fh.flush()

    def save(self):
        self.sync()

    def close(self):
        if self._fh is not None or self.vars:
            self.sync()
            self._fh.close()
            self._fh = None


def read_data_file(file_name):
    """Load every variable of a NetCDF-3 result file into a dict of arrays."""
    f = netcdf_file(file_name, 'r', mmap=False)
    out = {}
    for k, v in f.variables.items():
        a = np.array(v.data)
        out[k] = a.astype(a.dtype.newbyteorder('=')) if a.dtype.kind in 'fi' else a
    f.close()
    return out


def cg5_condensed():
    """
    Constants of the statically condensed CG5 L2 projection of a piecewise-constant field (du_smooth,
    diagnostic_equations.py:115-119) on a uniform periodic mesh, per unit cell length, for equispaced Lagrange
    nodes t_k = k/5 (the CG5 convention of this build):
        [sa, sb, cL, cR, e_1..4, l_1..4, r_1..4]
    vertex system  sb (v_{i-1} + v_{i+1}) + sa v_i = cL g_{i-1} + cR g_i,  interior dofs  w_k = e_k g_i + l_k v_i + r_k v_{i+1}.
    """
    from numpy.polynomial import polynomial as P
    t = np.arange(6) / 5.0
    basis = []
    for k in range(6):
        c = np.array([1.0])
        for j in range(6):
            if j != k:
                c = P.polymul(c, np.array([-t[j], 1.0]) / (t[k] - t[j]))
        basis.append(c)
    integ = lambda c: P.polyval(1.0, P.polyint(c)) - P.polyval(0.0, P.polyint(c))
    M = np.array([[integ(P.polymul(basis[a], basis[b])) for b in range(6)] for a in range(6)])
    c = np.array([integ(b) for b in basis])
    I, V = [1, 2, 3, 4], [0, 5]
    MII = M[np.ix_(I, I)]; MIV = M[np.ix_(I, V)]; MVI = M[np.ix_(V, I)]; MVV = M[np.ix_(V, V)]
    e = np.linalg.solve(MII, c[I])
    LR = -np.linalg.solve(MII, MIV)
    S = MVV - MVI @ np.linalg.solve(MII, MIV)
    cl = c[V] - MVI @ e
    return np.concatenate([[S[0, 0] + S[1, 1], S[0, 1], cl[1], cl[0]], e, LR[:, 0], LR[:, 1]])


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
