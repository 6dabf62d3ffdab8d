"""Regular grid <-> vector grid expansion (host side).

Same call signatures as the reference's Cython module (orbkit/cy_grid.pyx:14-97): x runs
slowest, z fastest.  These are O(N) host copies that only exist for API compatibility -- the
CUDA path never needs the expanded coordinates of a regular grid: the kernels derive
(x[i], y[j], z[k]) from the linear point index (0 input bytes per point).
"""
import math

import numpy


def _sin(a):
    """libm sin per element (numpy's SIMD sin may differ from the C library, which the reference calls, in the last bit)"""
    return numpy.array([math.sin(v) for v in a], dtype=numpy.float64)


def _cos(a):
    return numpy.array([math.cos(v) for v in a], dtype=numpy.float64)


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


def sph2cart(r, theta, phi):
    """(3, Nr*Ntheta*Nphi) Cartesian coordinates of a spherical product grid, r slowest, phi fastest, with the
    reference's expressions and multiplication order (cy_grid.pyx:58-75):
        x = r*sin(theta)*cos(phi),  y = r*sin(theta)*sin(phi),  z = r*cos(theta)"""
    r, theta, phi = (numpy.asarray(v, dtype=numpy.float64).reshape(-1) for v in (r, theta, phi))
    shape = (len(r), len(theta), len(phi))
    rs = r[:, None, None] * _sin(theta)[None, :, None]
    out = numpy.empty((3, len(r) * len(theta) * len(phi)), dtype=numpy.float64)
    out[0].reshape(shape)[...] = rs * _cos(phi)[None, None, :]
    out[1].reshape(shape)[...] = rs * _sin(phi)[None, None, :]
    out[2].reshape(shape)[...] = (r[:, None, None] * _cos(theta)[None, :, None]) * numpy.ones(len(phi))[None, None, :]
    return out


def cyl2cart(r, phi, zed):
    """(3, Nr*Nphi*Nzed) Cartesian coordinates of a cylindrical product grid (cy_grid.pyx:79-97):
        x = r*cos(phi),  y = r*sin(phi),  z = zed"""
    r, phi, zed = (numpy.asarray(v, dtype=numpy.float64).reshape(-1) for v in (r, phi, zed))
    shape = (len(r), len(phi), len(zed))
    out = numpy.empty((3, len(r) * len(phi) * len(zed)), dtype=numpy.float64)
    one = numpy.ones(len(zed))[None, None, :]
    out[0].reshape(shape)[...] = (r[:, None, None] * _cos(phi)[None, :, None]) * one
    out[1].reshape(shape)[...] = (r[:, None, None] * _sin(phi)[None, :, None]) * one
    out[2].reshape(shape)[...] = zed[None, None, :] * numpy.ones((len(r), len(phi), 1))
    return out
