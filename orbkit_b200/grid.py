"""Module-global grid state -- the second half of the input contract.

Mirror of the parts of orbkit/grid.py the hot path touches: the globals
`x, y, z, N_, min_, max_, delta_, d3r, is_vector, is_regular, is_initialized` (grid.py:663-677),
`grid_init` (:31-85), `set_grid` (:126), `grid2vector`/`vector2grid` (:205-243), `mv2g` (:268),
`adjust_to_geo` (:553), `reset_grid` (:651).  `rho_compute` mutates this state exactly like the
reference does (regular grid -> vector grid and back, core.py:433,569).

Device-generated product grids: `sph2cart_vector` (:373-395), `cyl2cart_vector` (:397-419),
`grid_sym_op` (:293-314) and `grid_translate` (:316-321) do not expand the coordinates on the host.  They record
a recipe (`_product`: kind, three axis vectors, affine map) from which the kernels generate the coordinates, and the
module attributes `x`, `y`, `z` are materialised lazily -- with the reference's own expressions -- the first time
somebody reads them (module-level __getattr__); from then on the grid is an ordinary vector grid.
`random_grid` (:421-456) and the file readers are host conveniences.
"""
import sys

import numpy

from . import cy_grid


_product = None     #: recipe of a device-generated product grid, or None


class ProductGrid:
    """(kind, axes, affine) of a spherical / cylindrical product grid whose Cartesian coordinates live only on the
    device until somebody asks for grid.x / grid.y / grid.z"""
    SPHERICAL, CYLINDRICAL = 2, 3

    def __init__(self, kind, a0, a1, a2):
        self.kind = kind
        self.axes = [numpy.array(a, dtype=numpy.float64).reshape(-1) for a in (a0, a1, a2)]
        self.matrix = None          # accumulated symmetry operations (3, 3)
        self.shift = None           # accumulated translation (3,)

    @property
    def npts(self):
        return len(self.axes[0]) * len(self.axes[1]) * len(self.axes[2])

    @property
    def affine(self):
        if self.matrix is None and self.shift is None:
            return None
        return (numpy.eye(3) if self.matrix is None else self.matrix,
                numpy.zeros(3) if self.shift is None else self.shift)

    def host_coordinates(self):
        """what the reference computes: cy_grid.sph2cart / cyl2cart, then numpy.dot(symop, xyz), then += shift"""
        f = cy_grid.sph2cart if self.kind == self.SPHERICAL else cy_grid.cyl2cart
        xyz = f(*self.axes)
        if self.matrix is not None:
            xyz = numpy.dot(self.matrix, xyz)
        if self.shift is not None:
            xyz = xyz + self.shift[:, None]
        return [numpy.ascontiguousarray(v) for v in xyz]


def _install_product(pg):
    """make `pg` the module grid; x, y, z become lazy"""
    global _product, is_initialized, is_vector, is_regular
    _product = pg
    g = globals()
    for c in ('x', 'y', 'z'):
        g.pop(c, None)
    is_initialized = True
    is_vector = True
    is_regular = False


def _ensure_host():
    """materialise x, y, z of a product grid (after this the grid is an ordinary vector grid)"""
    global _product, x, y, z
    if _product is not None:
        pg, _product = _product, None
        x, y, z = pg.host_coordinates()


def __getattr__(name):
    if name in ('x', 'y', 'z') and _product is not None:
        _ensure_host()
        return globals()[name]
    raise AttributeError('module %r has no attribute %r' % (__name__, name))


def product_grid():
    """the device recipe of the current grid, or None (ordinary regular / vector grid)"""
    return _product


def sph2cart_vector(r, theta, phi):
    """Spherical product grid (r, theta, phi) -> vector grid of Nr*Ntheta*Nphi points (grid.py:373-395);
    x = r sin(theta) cos(phi), y = r sin(theta) sin(phi), z = r cos(theta), generated on the device."""
    _install_product(ProductGrid(ProductGrid.SPHERICAL, r, theta, phi))


def cyl2cart_vector(r, phi, zed):
    """Cylindrical product grid (r, phi, zed) -> vector grid (grid.py:397-419); x = r cos(phi), y = r sin(phi)."""
    _install_product(ProductGrid(ProductGrid.CYLINDRICAL, r, phi, zed))


def grid_sym_op(symop):
    """Apply the symmetry operation `symop` (3x3) to the vector grid (grid.py:293-314)."""
    global x, y, z, is_regular
    symop = numpy.asarray(symop, dtype=numpy.float64)
    if symop.shape != (3, 3):
        raise ValueError('`symop` needs to be a numpy array with shape=(3,3)')
    if not is_initialized:
        raise ValueError('You have to initialize a grid before executing a symmetry operation on it. '
                         '(`grid.is_initialized == True`)')
    if _product is not None:
        _product.matrix = symop.copy() if _product.matrix is None else numpy.dot(symop, _product.matrix)
        if _product.shift is not None:
            _product.shift = numpy.dot(symop, _product.shift)
        return
    if not is_vector:
        grid2vector()
    x, y, z = numpy.dot(symop, numpy.array([x, y, z]))
    is_regular = False


def grid_translate(dx, dy, dz):
    """Translate the grid by (dx, dy, dz) (grid.py:316-321)."""
    global x, y, z
    if _product is not None:
        d = numpy.array([dx, dy, dz], dtype=numpy.float64)
        _product.shift = d if _product.shift is None else _product.shift + d
        return
    x += dx
    y += dy
    z += dz


def rot(ang, axis):
    """Rotation matrix about the x (0), y (1) or z (2) axis, angle in radians (grid.py:323-344)."""
    m = numpy.array([[numpy.cos(ang), numpy.sin(ang)], [-numpy.sin(ang), numpy.cos(ang)]])
    m = numpy.insert(numpy.insert(m, axis, 0, axis=0), axis, 0, axis=1)
    m[axis, axis] = 1
    return m


def reflect(plane):
    """Reflection matrix for the plane given by two axis indices, e.g. numpy.array([0,1]) = xy (grid.py:346-360)."""
    sigma = numpy.eye(3)
    sigma[3 - int(numpy.sum(plane)), 3 - int(numpy.sum(plane))] *= -1.0
    return sigma


def inversion():
    """Inversion matrix (grid.py:362-370)."""
    return -numpy.eye(3)


def random_grid(geo_spec, N=1e6, scale=0.5):
    """Normally distributed points around the atom positions (grid.py:421-456; like the reference the width is
    fixed at 0.5 whatever `scale` says)."""
    global x, y, z, is_initialized, is_vector, is_regular, _product
    geo_spec = numpy.array(geo_spec)
    N = int(N)
    pts = numpy.zeros((3, len(geo_spec), N))
    for d in range(3):
        for a in range(len(geo_spec)):
            pts[d, a, :] = numpy.random.normal(loc=geo_spec[a, d], scale=0.5, size=N)
    pts = pts.reshape((3, N * len(geo_spec)))
    _product = None
    x, y, z = pts[0], pts[1], pts[2]
    is_initialized = True
    is_vector = True
    is_regular = False


def grid_init(is_vector=False, force=False):
    """Set up x, y, z from min_/max_/N_ (or delta_) (grid.py:31-85)."""
    global x, y, z, d3r, min_, max_, N_, delta_, is_initialized, is_regular, _product
    if is_initialized and not force:
        return 0
    _product = None
    axes = [None, None, None]
    for i in range(3):
        if max_[i] == min_[i]:
            axes[i] = numpy.array([min_[i]], dtype=numpy.float64)
            delta_[i] = 1
            N_[i] = 1
        elif delta_[i]:
            step = float(numpy.asarray(delta_[i]).reshape(-1)[0])
            axes[i] = numpy.arange(min_[i], max_[i] + step, step, dtype=numpy.float64)
            delta_[i] = step
            N_[i] = len(axes[i])
        else:
            axes[i] = numpy.array(numpy.linspace(min_[i], max_[i], N_[i]), dtype=numpy.float64)
            delta_[i] = axes[i][1] - axes[i][0]
    x, y, z = axes
    d3r = float(numpy.prod([float(numpy.asarray(d).reshape(-1)[0]) for d in delta_]))
    is_initialized = True
    is_regular = True
    if is_vector:
        grid2vector()
    else:
        setattr(sys.modules[__name__], 'is_vector', False)


init = grid_init
init_grid = grid_init


def get_grid(start='\t'):
    _ensure_host()
    out = ''
    for c, g, i in (('x', x, 0), ('y', y, 1), ('z', z, 2)):
        out += '%s%s[0] = %.2f %s[-1] = %.2f N%s = %d ' % (start, c, g[0], c, g[-1], c, len(g))
        if max_[i] != min_[i] and delta_[i] != 0.:
            out += 'd%s = %.3f' % (c, delta_[i])
        out += '\n'
    return out


def tolist():
    _ensure_host()
    return [numpy.copy(x), numpy.copy(y), numpy.copy(z)]


def todict():
    _ensure_host()
    return {'x': x, 'y': y, 'z': z}


def get_shape():
    _ensure_host()
    if not is_initialized:
        raise ValueError('`grid.get_shape` requires the grid to be initialized.')
    return (len(x),) if is_vector else tuple(N_)


def set_grid(xnew, ynew, znew, is_vector):
    """Install user coordinates (grid.py:126-184)."""
    global x, y, z, is_initialized, is_regular
    reset_grid()
    axes = []
    for i, c in enumerate((xnew, ynew, znew)):
        if isinstance(c, (int, float)):
            c = numpy.array([c], dtype=numpy.float64)
        elif isinstance(c, (list, tuple)):
            c = numpy.array(c, dtype=numpy.float64)
        elif not isinstance(c, numpy.ndarray):
            raise TypeError('%s (dimension %d) is of inappropriate type. (%s)' % ('xyz'[i], i, type(c)))
        axes.append(numpy.asarray(c, dtype=numpy.float64).reshape((-1,)))
    x, y, z = axes
    is_initialized = True
    if isinstance(is_vector, bool):
        setattr(sys.modules[__name__], 'is_vector', is_vector)
    is_regular = (is_vector == False)  # noqa: E712  (reference semantics: None -> False)
    set_boundaries(is_regular)
    return 'Grid has been set up (%d x %d x %d).' % (len(x), len(y), len(z))


def set_boundaries(is_regular, Nx=None, Ny=None, Nz=None):
    global min_, max_, delta_, N_
    _ensure_host()
    min_ = [v.min() if len(v) else 0.0 for v in (x, y, z)]
    max_ = [v.max() if len(v) else 0.0 for v in (x, y, z)]
    N_ = [len(x), len(y), len(z)]
    if is_regular:
        f = lambda v: 1.0 if len(v) <= 1 else v[1] - v[0]
        delta_ = [f(x), f(y), f(z)]
    elif all([Nx, Ny, Nz]):
        N_ = [Nx, Ny, Nz]
        g = numpy.array([x, y, z]).reshape(3, Nx, Ny, Nz)
        delta_ = [g[0, 1, 0, 0] - g[0, 0, 0, 0] if Nx > 1 else 1.0,
                  g[1, 0, 1, 0] - g[1, 0, 0, 0] if Ny > 1 else 1.0,
                  g[2, 0, 0, 1] - g[2, 0, 0, 0] if Nz > 1 else 1.0]


def get_bbox():
    _ensure_host()
    bbox = numpy.zeros(6)
    bbox[::2] = min_
    bbox[1::2] = max_
    return bbox


def grid2vector():
    """Regular (x,y,z axes) -> (3, Nx*Ny*Nz) vector grid, in place (grid.py:205-219)."""
    global x, y, z, is_vector, is_regular
    if not is_initialized:
        raise ValueError('You have to initialize a grid before calling `grid.grid2vector`.')
    _ensure_host()
    x, y, z = cy_grid.grid2vector(x, y, z)
    is_vector = True
    is_regular = True


def vector2grid(Nx, Ny, Nz):
    """Inverse of grid2vector (grid.py:221-243)."""
    global x, y, z, is_vector
    if not is_initialized:
        raise ValueError('You have to initialize a grid before calling `grid.vector2grid`.')
    _ensure_host()
    if not is_regular:
        raise ValueError('The grid has to regular. (`grid.is_regular == True`)')
    if not (len(x) == len(y) == len(z)):
        raise ValueError('Not a valid vector grid, i.e., dimensions of x-, y-, and z- coordinate differ.')
    if (Nx * Ny * Nz) != len(x):
        raise ValueError('It has to hold that `len(x) = (N_x * N_y * N_z)`')
    x, y, z = cy_grid.vector2grid(x, y, z, Nx, Ny, Nz)
    is_vector = False


def matrix_grid2vector(matrix):
    matrix = numpy.asarray(matrix, dtype=float)
    if matrix.ndim != 3:
        raise ValueError('`matrix` has to be 3d matrix.')
    return numpy.reshape(matrix, (-1,))


def matrix_vector2grid(matrix, Nx=None, Ny=None, Nz=None):
    matrix = numpy.asarray(matrix, dtype=float)
    if matrix.ndim != 1 or (Nx * Ny * Nz) != len(matrix):
        raise ValueError('`matrix` has to be one dimensional with the length N_x * N_y * N_z.')
    return numpy.reshape(matrix, (Nx, Ny, Nz))


def mv2g(**kwargs):
    """(..., Nx*Ny*Nz, ...) -> (..., Nx, Ny, Nz, ...) using the global N_ (grid.py:268-291)."""
    out = {}
    n = int(numpy.prod(N_))
    for key, val in kwargs.items():
        val = numpy.asarray(val, dtype=float)
        where = [i for i, s in enumerate(val.shape) if s == n][0]
        out[key] = val.reshape(val.shape[:where] + tuple(N_) + val.shape[where + 1:])
    return list(out.values())[0] if len(out) == 1 else out


def adjust_to_geo(qc, extend=5.0, step=0.1):
    """Fit the regular-grid box to the molecule (grid.py:553-582)."""
    global min_, max_, N_, delta_, is_vector, is_initialized
    geo = numpy.asarray(qc.geo_spec, dtype=float)
    for i in range(3):
        min_[i] = min(geo[:, i]) - extend
        max_[i] = max(geo[:, i]) + extend
        dist = max_[i] - min_[i]
        N_[i] = int(numpy.ceil(dist / step)) + 1
        rest = (N_[i] - 1) * step - dist
        min_[i] -= rest / 2.
        max_[i] += rest / 2.
        delta_[i] = step
    is_vector = False
    is_initialized = False


def reset_grid():
    global is_initialized, is_vector, is_regular, min_, max_, N_, delta_, _product, x, y, z
    if _product is not None:             # drop a device recipe; x, y, z fall back to the defaults
        _product = None
        x, y, z = numpy.array([0.0]), numpy.array([0.0]), numpy.array([0.0])
    is_initialized = False
    is_vector = True
    is_regular = False
    min_ = [-8.0, -8.0, -8.0]
    max_ = [8.0, 8.0, 8.0]
    N_ = [101, 101, 101]
    delta_ = numpy.zeros((3, 1))


min_ = [-8.0, -8.0, -8.0]    #: minimum grid values (regular grid)
max_ = [8.0, 8.0, 8.0]       #: maximum grid values (regular grid)
N_ = [101, 101, 101]         #: number of grid points (regular grid)
x = numpy.array([0.0])
y = numpy.array([0.0])
z = numpy.array([0.0])
delta_ = numpy.zeros((3, 1))  #: grid spacing
d3r = 0.0                     #: volume element
is_initialized = False
is_vector = True
is_regular = False
