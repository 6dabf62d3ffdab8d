"""Regular grid <-> vector grid expansion (host side).

Same call signatures as the reference's Cython module (orbkit/cy_grid.pyx:14-55): x runs
slowest, z fastest.  These are O(N) host copies that only exist for API compatibility -- the
CUDA path never needs the expanded coordinates of a regular grid: the kernels derive
(x[i], y[j], z[k]) from the linear point index (0 input bytes per point).
"""
import numpy


def grid2vector(x, y, z):
    x, y, z = (numpy.asarray(v, dtype=numpy.float64) for v in (x, y, z))
    nx, ny, nz = len(x), len(y), len(z)
    out = numpy.empty((3, nx * ny * nz), dtype=numpy.float64)
    out[0].reshape(nx, ny, nz)[...] = x[:, None, None]
    out[1].reshape(nx, ny, nz)[...] = y[None, :, None]
    out[2].reshape(nx, ny, nz)[...] = z[None, None, :]
    return out


def vector2grid(x, y, z, Nx, Ny, Nz):
    x, y, z = (numpy.asarray(v, dtype=numpy.float64) for v in (x, y, z))
    return (x[::Ny * Nz][:Nx].copy(), y[::Nz][:Ny].copy(), z[:Nz].copy())
