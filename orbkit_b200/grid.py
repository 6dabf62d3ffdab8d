"""Module-global grid state -- the second half of the input contract.

Mirror of the parts of orbkit/grid.py the hot path touches: the globals
`x, y, z, N_, min_, max_, delta_, d3r, is_vector, is_regular, is_initialized` (grid.py:663-677),
`grid_init` (:31-85), `set_grid` (:126), `grid2vector`/`vector2grid` (:205-243), `mv2g` (:268),
`adjust_to_geo` (:553), `reset_grid` (:651).  `rho_compute` mutates this state exactly like the
reference does (regular grid -> vector grid and back, core.py:433,569).  Symmetry operations,
file readers and random/spherical grids are host conveniences outside the hot path.
"""
import sys

import numpy

from . import cy_grid


def grid_init(is_vector=False, force=False):
    """Set up x, y, z from min_/max_/N_ (or delta_) (grid.py:31-85)."""
    global x, y, z, d3r, min_, max_, N_, delta_, is_initialized, is_regular
    if is_initialized and not force:
        return 0
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
    out = ''
    for c, g, i in (('x', x, 0), ('y', y, 1), ('z', z, 2)):
        out += '%s%s[0] = %.2f %s[-1] = %.2f N%s = %d ' % (start, c, g[0], c, g[-1], c, len(g))
        if max_[i] != min_[i] and delta_[i] != 0.:
            out += 'd%s = %.3f' % (c, delta_[i])
        out += '\n'
    return out


def tolist():
    return [numpy.copy(x), numpy.copy(y), numpy.copy(z)]


def todict():
    return {'x': x, 'y': y, 'z': z}


def get_shape():
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
    bbox = numpy.zeros(6)
    bbox[::2] = min_
    bbox[1::2] = max_
    return bbox


def grid2vector():
    """Regular (x,y,z axes) -> (3, Nx*Ny*Nz) vector grid, in place (grid.py:205-219)."""
    global x, y, z, is_vector, is_regular
    if not is_initialized:
        raise ValueError('You have to initialize a grid before calling `grid.grid2vector`.')
    x, y, z = cy_grid.grid2vector(x, y, z)
    is_vector = True
    is_regular = True


def vector2grid(Nx, Ny, Nz):
    """Inverse of grid2vector (grid.py:221-243)."""
    global x, y, z, is_vector
    if not is_initialized:
        raise ValueError('You have to initialize a grid before calling `grid.vector2grid`.')
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
    global is_initialized, is_vector, is_regular, min_, max_, N_, delta_
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
