"""Host-side helpers of the hot path: Cartesian exponent orderings, the Cartesian -> real
spherical table, `drv` validation and dtype coercion.

Mirrors the symbols the grid path uses from orbkit/tools.py (exp :118-135, exp_wfn :138-145,
cart2sph :155-191, get_cart2sph :193, validate_drv :225-239, require :290, convert :297,
zeros :303, reshape :309).  Everything else in that module (NIST masses, plotting) is out of
scope.
"""
import string

import numpy

# angular momentum letter -> l  (s p d f g h i k ...)
orbit = 'spd' + string.ascii_lowercase[5:].replace('s', '').replace('p', '')
lquant = dict((j, i) for i, j in enumerate(orbit))


def l_deg(l=0, ao=None, cartesian_basis=True):
    """Number of functions of a shell (tools.py:86-113)."""
    if ao is not None:
        if ao == 's':
            return 1
        l = len(ao)
    elif not isinstance(l, (int, numpy.integer)):
        l = lquant[l]
    return (l + 1) * (l + 2) // 2 if cartesian_basis else 2 * l + 1


def _molden_order(l):
    table = {
        0: ['', ],
        1: ['x', 'y', 'z'],
        2: ['xx', 'yy', 'zz', 'xy', 'xz', 'yz'],
        3: ['xxx', 'yyy', 'zzz', 'xyy', 'xxy', 'xxz', 'xzz', 'yzz', 'yyz', 'xyz'],
        4: ['xxxx', 'yyyy', 'zzzz', 'xxxy', 'xxxz', 'xyyy', 'yyyz', 'xzzz', 'yzzz',
            'xxyy', 'xxzz', 'yyzz', 'xxyz', 'xyyz', 'xyzz'],
    }[l]
    return [(s.count('x'), s.count('y'), s.count('z')) for s in table]


#: Molden order of the Cartesian exponents (lx,ly,lz) per l (tools.py:118-135)
exp = [_molden_order(l) for l in range(5)]
#: wfn order: differs from Molden only for f (tools.py:138-145)
exp_wfn = exp[:3] + [[(s.count('x'), s.count('y'), s.count('z')) for s in
                      ['xxx', 'yyy', 'zzz', 'xxy', 'xxz', 'yyz', 'xyy', 'xzz', 'yzz', 'xyz']],
                     exp[4]]

_s = numpy.sqrt


def _t(terms, factor=1.):
    """'xxz:-1.5 ...' -> [[(lx,ly,lz)...], [coef...], factor]"""
    e, c = [], []
    for mono, coef in terms:
        e.append((mono.count('x'), mono.count('y'), mono.count('z')))
        c.append(float(coef))
    return [e, c, float(factor)]


#: cart2sph[l][l+m] = [exponent triples, coefficients, global factor]; same content and
#: term order as the reference table (tools.py:155-191, after Schlegel & Frisch, IJQC 54, 83
#: (1995)), INCLUDING its two wrong g rows ((4,-1) names yyyz twice; (4,0) is mis-scaled) so
#: that parity mode reproduces the reference.  `cart2sph_g_fixed` holds the corrected l=4 rows.
cart2sph = [
    [_t([('', 1.)])],
    [_t([('y', 1.)]), _t([('z', 1.)]), _t([('x', 1.)])],
    [_t([('xy', 1.)]),
     _t([('yz', 1.)]),
     _t([('zz', 1.), ('xx', -1 / 2.), ('yy', -1 / 2.)]),
     _t([('xz', 1.)]),
     _t([('xx', 1.), ('yy', -1.)], _s(3) / 2.)],
    [_t([('yyy', -_s(5)), ('xxy', 3.)], 1 / (2. * _s(2))),
     _t([('xyz', 1.)]),
     _t([('yzz', _s(3 / 5.)), ('yyy', -_s(3) / 4.), ('xxy', -_s(3) / (4. * _s(5)))], _s(2)),
     _t([('zzz', 1.), ('xxz', -3 / (2 * _s(5))), ('yyz', -3 / (2 * _s(5)))]),
     _t([('xzz', _s(3 / 5.)), ('xxx', -_s(3) / 4.), ('xyy', -_s(3) / (4. * _s(5)))], _s(2)),
     _t([('xxz', 1.), ('yyz', -1.)], _s(3) / 2.),
     _t([('xxx', _s(5)), ('xyy', -3.)], 1 / (2. * _s(2)))],
    [_t([('xxxy', 1.), ('xyyy', -1.)], _s(2) * _s(5 / 8.)),
     _t([('yyyz', -_s(5) / 4.), ('xxyz', 3 / 4.)], _s(2)),
     _t([('xyzz', 3 / _s(14)), ('xxxy', -_s(5) / (2 * _s(14))), ('xyyy', -_s(5) / (2 * _s(14)))], _s(2)),
     _t([('yyyz', _s(5 / 7.)), ('yyyz', -3 * _s(5) / (4. * _s(7))), ('xxyz', -3 / (4. * _s(7)))], _s(2)),
     _t([('zzzz', 1.), ('xxxx', 3 / 8.), ('yyyy', 3 / 8.), ('xxzz', -3 * _s(3) / _s(35)),
         ('yyzz', -3 * _s(3) / _s(35)), ('xxyy', -1 / 4.)], _s(2)),
     _t([('xzzz', _s(5 / 7.)), ('xxxz', -3 * _s(5) / (4. * _s(7))), ('xyyz', -3 / (4. * _s(7)))], _s(2)),
     _t([('xxzz', 3 * _s(3) / (2. * _s(14))), ('yyzz', -3 * _s(3) / (2. * _s(14))),
         ('xxxx', -_s(5) / (4. * _s(2))), ('yyyy', _s(5) / (4. * _s(2)))], _s(2)),
     _t([('xxxz', _s(5) / 4.), ('xyyz', -3 / 4.)], _s(2)),
     _t([('xxxx', _s(35) / (8. * _s(2))), ('yyyy', _s(35) / (8. * _s(2))),
         ('xxyy', -3 * _s(3) / (4. * _s(2)))], _s(2))],
]


def get_cart2sph(l, m):
    """Linear combination of Cartesian Gaussians giving the real spherical (l,m) (tools.py:193)."""
    return cart2sph[l][l + m]


_DRV = {None: 0, 'None': 0, '': 0, 'x': 1, 'y': 2, 'z': 3, 'xx': 4, 'x2': 4, 'yy': 5, 'y2': 5,
        'zz': 6, 'z2': 6, 'xy': 7, 'yx': 7, 'xz': 8, 'zx': 8, 'yz': 9, 'zy': 9}


def validate_drv(drv):
    """drv string -> derivative code 0..9; ints 0..9 pass through (tools.py:225-239)."""
    if drv is None or isinstance(drv, str):
        if drv not in _DRV:
            raise ValueError("The selection `drv=%s` is not valid!" % drv)
        return _DRV[drv]
    if isinstance(drv, (int, numpy.integer)) and not isinstance(drv, bool) and 0 <= drv <= 9:
        return int(drv)
    raise ValueError("The selection `drv=%s` is not valid!" % drv)


def require(data, dtype='f', requirements='CA'):
    """C-contiguous aligned float64 / intc array (tools.py:290-295)."""
    if dtype == 'f':
        dtype = numpy.float64
    elif dtype == 'i':
        dtype = numpy.intc
    return numpy.require(data, dtype=dtype, requirements='CA')


def convert(data, was_vector, N):
    data = numpy.array(data, order='C')
    if not was_vector:
        data = data.reshape(data.shape[:-1] + tuple(N), order='C')
    return data


def zeros(shape, name, hdf5_file=None, chunks=True):
    if hdf5_file is None:
        return numpy.zeros(shape)
    return hdf5_file.create_dataset(name, shape, dtype=numpy.float64, chunks=chunks)


def reshape(data, shape, save_hdf5=False):
    if not save_hdf5:
        return data.reshape(shape)
    data.attrs['shape'] = shape
    return data[...].reshape(shape)
