"""ctypes binding of libokb200.so (include/okb200.h) -- the only door to the CUDA kernels.

There is NO CPU fallback: if the shared library is missing or no sm_100 device is present every
compute entry point raises.  The library is built in-tree by `__graft_entry__.build()` /
`orbkit_b200._lib.build()` with
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -Xcompiler -fPIC -c  (one object per csrc/*.cu,
    compiled in parallel) and nvcc -shared to link
"""
import ctypes
import os
import re
import subprocess
import threading

import numpy

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('OKB_LIB_PATH') or os.path.join(_HERE, 'libokb200.so')   # override: A/B builds only
SRC = os.path.join(_HERE, 'csrc', 'okb200.cu')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC']

OKB_FLAG_EXACT_MIXED = 1
OKB_FLAG_OUT_DEVICE = 2
OKB_FLAG_IN_DEVICE = 4
OKB_FLAG_CI_FAST = 8
OKB_CI_RHO, OKB_CI_JAB, OKB_CI_A_NABLA_B, OKB_CI_PAIRS, OKB_CI_JAB_PAIRS = 0, 1, 2, 3, 4

c_int_p = ctypes.POINTER(ctypes.c_int)
c_double_p = ctypes.POINTER(ctypes.c_double)
c_void_pp = ctypes.POINTER(ctypes.c_void_p)
ll = ctypes.c_longlong

# name -> (restype, argtypes); every symbol include/okb200.h declares
SIGNATURES = {
    'okb_last_error': (ctypes.c_char_p, []),
    'okb_version': (ctypes.c_int, []),
    'okb_device_count': (ctypes.c_int, [c_int_p]),
    'okb_ctx_create': (ctypes.c_int, [ctypes.c_int, c_void_pp]),
    'okb_ctx_destroy': (ctypes.c_int, [ctypes.c_void_p]),
    'okb_ctx_sync': (ctypes.c_int, [ctypes.c_void_p]),
    'okb_ctx_stream': (ctypes.c_void_p, [ctypes.c_void_p]),
    'okb_ctx_launch_count': (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ll)]),
    'okb_ctx_traffic': (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ll), ctypes.POINTER(ll)]),
    'okb_ctx_last_kernel': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int]),
    'okb_measure_fp64': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, c_double_p, c_double_p]),
    'okb_aocreator': (ctypes.c_int, [ctypes.c_void_p, c_int_p, c_int_p, c_double_p, c_int_p, c_double_p,
                                     c_int_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                     c_double_p, c_double_p, c_double_p, ll, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_uint, c_double_p]),
    'okb_lcreator': (ctypes.c_int, [ctypes.c_void_p, c_double_p, ll, c_int_p, c_double_p, c_double_p,
                                    c_double_p, c_double_p, c_double_p, ll, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_int, ctypes.c_int, ctypes.c_uint]),
    'okb_mocreator': (ctypes.c_int, [ctypes.c_void_p, c_double_p, c_double_p, ctypes.c_int, ll,
                                     ctypes.c_int, c_double_p]),
    'okb_aonorm': (ctypes.c_double, [ctypes.c_int] * 3 + [ctypes.c_double, ctypes.c_int]),
    'okb_aoxyz': (ctypes.c_double, [ctypes.c_double] * 3 + [ctypes.c_int] * 3 + [ctypes.c_double, ctypes.c_int]),
    'okb_basis_create': (ctypes.c_int, [ctypes.c_void_p, c_int_p, c_int_p, c_double_p, c_int_p, c_double_p,
                                        c_int_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, c_double_p, c_void_pp]),
    'okb_basis_set_cart2sph': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_int_p, c_int_p, c_double_p]),
    'okb_basis_info': (ctypes.c_int, [ctypes.c_void_p, c_int_p, c_int_p, c_int_p, c_int_p]),
    'okb_basis_destroy': (ctypes.c_int, [ctypes.c_void_p]),
    'okb_mo_create': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, c_double_p, c_double_p,
                                     c_void_pp]),
    'okb_mo_destroy': (ctypes.c_int, [ctypes.c_void_p]),
    'okb_grid_regular': (ctypes.c_int, [ctypes.c_void_p, c_double_p, ctypes.c_int, c_double_p, ctypes.c_int,
                                        c_double_p, ctypes.c_int, c_void_pp]),
    'okb_grid_vector': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ll,
                                       ctypes.c_int, c_void_pp]),
    'okb_grid_product': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_double_p, ctypes.c_int, c_double_p,
                                        ctypes.c_int, c_double_p, ctypes.c_int, c_double_p, c_void_pp]),
    'okb_grid_size': (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ll)]),
    'okb_grid_destroy': (ctypes.c_int, [ctypes.c_void_p]),
    'okb_eval_ao': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ll, ll, c_int_p,
                                   ctypes.c_int, ctypes.c_void_p, ctypes.c_uint]),
    'okb_eval_mo': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ll, ll, c_int_p,
                                   ctypes.c_int, ctypes.c_void_p, ctypes.c_uint]),
    'okb_eval_rho': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ll, ll, c_int_p,
                                    ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.c_uint]),
    'okb_eval_ao_ld': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ll, ll, c_int_p,
                                      ctypes.c_int, ctypes.c_void_p, ll, ctypes.c_uint]),
    'okb_eval_mo_ld': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ll, ll, c_int_p,
                                      ctypes.c_int, ctypes.c_void_p, ll, ctypes.c_uint]),
    'okb_eval_rho_ld': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ll, ll, c_int_p,
                                       ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ll, ctypes.c_void_p,
                                       ctypes.c_uint]),
    'okb_ci_td': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ll, c_double_p, ctypes.c_void_p, ll,
                                 ctypes.c_void_p, ll, ctypes.c_uint]),
    'okb_ci_jab_full': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ll, ll, c_double_p,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, ctypes.c_void_p, ll,
                                       ctypes.c_uint]),
    'okb_aooverlap': (ctypes.c_int, [ctypes.c_void_p, c_double_p, c_double_p, ctypes.c_int, c_int_p, c_int_p,
                                     ctypes.c_int, c_int_p, c_double_p, c_int_p, c_int_p, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_int, c_double_p]),
    'okb_ci_contract': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ll, ll, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_int, c_double_p, c_int_p, c_int_p,
                                       ctypes.c_void_p, ll, ctypes.c_uint]),
    'okb_eval_ci': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ll, ll, ctypes.c_int,
                                   c_int_p, ctypes.c_int, c_double_p, c_int_p, c_int_p, ctypes.c_void_p,
                                   ctypes.c_uint]),
    'okb_cube_body_bytes': (ll, [ctypes.c_int, ll, ll, ll]),
    'okb_format_cube': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ll, ll, ll, ctypes.c_void_p, ll,
                                       ctypes.c_uint]),
}

_lock = threading.Lock()
_lib = None


class OkbError(RuntimeError):
    pass


def build(force=False, verbose=False, jobs=None):
    """Compile csrc/*.cu for sm_100a into orbkit_b200/libokb200.so (works without a GPU).

    One object per translation unit (okb200.cu = host side + small kernels, inst_*.cu = groups of
    kernel-template instantiations), compiled in parallel, then linked by nvcc."""
    from concurrent.futures import ThreadPoolExecutor
    csrc = os.path.join(_HERE, 'csrc')
    headers = [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(('.cuh', '.h'))]
    headers.append(os.path.join(os.path.dirname(_HERE), 'include', 'okb200.h'))
    units = sorted(os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith('.cu'))
    objdir = os.path.join(csrc, 'build')
    os.makedirs(objdir, exist_ok=True)
    hdr_time = max(os.path.getmtime(h) for h in headers)
    inc_re = re.compile(r'^\s*#\s*include\s+"([^"]+)"', re.M)

    def deps_time(path, seen):
        """newest modification time of `path` and the project headers it includes (recursively); a header that
        cannot be found next to the sources counts as 'any header changed'"""
        if path in seen:
            return 0.0
        seen.add(path)
        t = os.path.getmtime(path)
        with open(path) as f:
            text = f.read()
        for name in inc_re.findall(text):
            for d in (os.path.dirname(path), csrc, os.path.join(os.path.dirname(_HERE), 'include')):
                cand = os.path.join(d, name)
                if os.path.exists(cand):
                    t = max(t, deps_time(cand, seen))
                    break
            else:
                t = max(t, hdr_time)
        return t
    nvcc = os.environ.get('NVCC', 'nvcc')
    extra = os.environ.get('OKB_NVCC_EXTRA', '').split()        # A/B builds, e.g. -DOKB_REFILL_DEP

    def obj_of(u):
        return os.path.join(objdir, os.path.basename(u)[:-3] + '.o')

    def stale(u):
        o = obj_of(u)
        return force or not os.path.exists(o) or os.path.getmtime(o) < deps_time(u, set())

    todo = [u for u in units if stale(u)]

    def compile_one(u):
        cmd = [nvcc] + NVCC_FLAGS + extra + ['-c', '-o', obj_of(u), u]
        if verbose:
            print(' '.join(cmd))
        subprocess.check_call(cmd)

    if todo:
        with ThreadPoolExecutor(max_workers=jobs or min(len(todo), os.cpu_count() or 4)) as pool:
            list(pool.map(compile_one, todo))
    objs = [obj_of(u) for u in units]
    if todo or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(o) for o in objs):
        cmd = [nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-o', LIB_PATH] + objs
        if verbose:
            print(' '.join(cmd))
        subprocess.check_call(cmd)
    return LIB_PATH


def load():
    """Load libokb200.so; raise (loudly) if it has not been built."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise OkbError('orbkit_b200: %s is missing -- run `python -c "import __graft_entry__ as g; '
                           'g.build()"` (nvcc, sm_100a). There is no CPU fallback.' % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)      # AttributeError if the library lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def check(status):
    if status != 0:
        msg = load().okb_last_error().decode('utf-8', 'replace')
        if status == 1:
            raise ValueError(msg)
        if status == 3:
            raise MemoryError(msg)
        raise OkbError(msg)


def dptr(a):
    return a.ctypes.data_as(c_double_p)


def iptr(a):
    return a.ctypes.data_as(c_int_p)


def f64(a):
    return numpy.require(a, dtype=numpy.float64, requirements='CA')


def i32(a):
    return numpy.require(a, dtype=numpy.intc, requirements='CA')
