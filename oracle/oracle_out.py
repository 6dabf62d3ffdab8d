"""CPU ORACLE for the cube output sink (orbkit/output/cube.py:5-101).

TEST INFRASTRUCTURE ONLY (see oracle.py): imported by tests/ and bench legs as the checker, never by orbkit_b200.

    cube_text(data, geo_info, geo_spec, min_, N_, delta_, comments='', labels=None) -> str

restates cube_creator line by line (same string operations, same loop order), with the grid attributes the reference
reads from its module globals passed explicitly.  Pinned byte for byte to files written by the reference's own
cube_creator (tests/golden/cube_text.npz, generator tests/golden/make_golden_cube.py), including the special values of
'%.5E' (exact ties, three-digit exponents, signed zeros, subnormals, inf, nan).
"""
import numpy as np


def cube_text(data, geo_info, geo_spec, min_, N_, delta_, comments='', labels=None):
    data = np.array(data)
    if data.ndim < 3:
        raise AssertionError('data.ndim < ndim of grid')
    elif data.ndim == 3:
        data = data[np.newaxis]
    elif data.ndim > 4:
        raise AssertionError('data.ndim > (ndim of grid) +2')
    if labels is not None:                                               # cube.py:32-39
        if labels is True or labels == 'auto':
            labels = list(range(len(data)))
        assert len(labels) == len(data)
        labels = [int(j) for j in labels]
    assert data.shape[1:] == tuple(N_), 'The grid does not fit the data.'
    s = 'orbkit calculation\n'                                           # cube.py:47-84
    s += ' %(f)s\n' % {'f': comments}
    s += ('%(at)d' % {'at': (-1) ** (labels is not None) * len(geo_info)}).rjust(5)
    for ii in range(3):
        s += ('%(min)0.6f' % {'min': min_[ii]}).rjust(12)
    if len(data) > 1:
        s += ('%d' % len(data)).rjust(12)
    for ii in range(3):
        s += '\n'
        s += ('%(N)d' % {'N': N_[ii]}).rjust(5)
        for jj in range(3):
            s += ('%(dr)0.6f' % {'dr': delta_[ii] if jj == ii else 0}).rjust(12)
    s += '\n'
    for ii in range(len(geo_info)):
        s += ('%(N)d' % {'N': round(float(geo_info[ii][2]))}).rjust(5)
        s += ('%(ch)0.6f' % {'ch': float(geo_info[ii][1])}).rjust(12)
        for jj in range(3):
            s += ('%(r)0.6f' % {'r': geo_spec[ii][jj]}).rjust(12)
        s += '\n'
    if labels is not None:
        s += ('%(N)d' % {'N': len(data)}).rjust(5)
        c = 0
        for j in labels:
            c += 1
            s += str(j).rjust(5)
            if c % 9 == 8:
                s += '\n'
        s += '\n'
    parts = [s]                                                          # cube.py:86-96
    for rr in range(data.shape[1]):
        for ss in range(data.shape[2]):
            c = 0
            row = ''
            for tt in range(data.shape[3]):
                for dd in data[:, rr, ss, tt]:
                    row += ('%(data).5E' % {'data': dd}).rjust(13)
                    if c % 6 == 5:
                        row += '\n'
                    c += 1
            row += '\n'
            parts.append(row)
    return ''.join(parts)


def cube_body(data):
    """only the data block (bytes), for parity checks of the device formatter"""
    data = np.array(data)
    if data.ndim == 3:
        data = data[np.newaxis]
    parts = []
    for rr in range(data.shape[1]):
        for ss in range(data.shape[2]):
            c = 0
            row = ''
            for tt in range(data.shape[3]):
                for dd in data[:, rr, ss, tt]:
                    row += ('%(data).5E' % {'data': dd}).rjust(13)
                    if c % 6 == 5:
                        row += '\n'
                    c += 1
            row += '\n'
            parts.append(row)
    return ''.join(parts).encode('utf-8')
