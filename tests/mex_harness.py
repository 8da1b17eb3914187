"""ctypes driver for tests/mex_stub/libgpz_mex_harness.so: matlab/gpz_b200_mex.cpp linked against the minimal libmx mock
(tests/mex_stub/mex_mock.cpp).  `call("eval", h, theta, nlhs=3)` does what `[f,g,stats] = gpz_b200_mex('eval',h,theta)` does
in MATLAB: NumPy arrays become column-major mxArrays, dicts become 1 x 1 structs, str becomes char, bool arrays logical,
None becomes []; a raised mexErrMsgIdAndTxt becomes MexError(id, message)."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(HERE, "mex_stub", "libgpz_mex_harness.so")
_lib = None


class MexError(RuntimeError):
    def __init__(self, ident, message):
        super().__init__(f"{ident}: {message}")
        self.ident = ident
        self.message = message


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(PATH)
        vp = C.c_void_p
        lib.mock_double.restype = vp
        lib.mock_double.argtypes = [C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_int64)]
        lib.mock_logical.restype = vp
        lib.mock_logical.argtypes = [C.POINTER(C.c_ubyte), C.c_int64]
        lib.mock_string.restype = vp
        lib.mock_string.argtypes = [C.c_char_p]
        lib.mock_struct.restype = vp
        lib.mock_empty.restype = vp
        lib.mock_set_field.argtypes = [vp, C.c_char_p, vp]
        lib.mock_free.argtypes = [vp]
        lib.mock_ndim.argtypes = [vp]
        lib.mock_dim.restype = C.c_int64
        lib.mock_dim.argtypes = [vp, C.c_int]
        lib.mock_data.restype = C.POINTER(C.c_double)
        lib.mock_data.argtypes = [vp]
        lib.mock_printed.restype = C.c_char_p
        lib.mock_call.argtypes = [C.c_int, C.POINTER(vp), C.c_int, C.POINTER(vp), C.c_char_p, C.c_int]
        _lib = lib
    return _lib


def to_mx(v):
    lib = load()
    if v is None:
        return lib.mock_empty()
    if isinstance(v, str):
        return lib.mock_string(v.encode())
    if isinstance(v, dict):
        s = lib.mock_struct()
        for k, x in v.items():
            lib.mock_set_field(s, k.encode(), to_mx(x))
        return s
    a = np.asarray(v)
    if a.dtype == np.bool_:
        a = np.ascontiguousarray(a.reshape(-1).astype(np.uint8))
        return lib.mock_logical(a.ctypes.data_as(C.POINTER(C.c_ubyte)), a.size)
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 0:
        a = a.reshape(1, 1)
    if a.ndim == 1:
        a = a.reshape(-1, 1)                      # MATLAB column vector
    dims = (C.c_int64 * a.ndim)(*a.shape)
    flat = np.ascontiguousarray(a.reshape(-1, order="F"))
    return lib.mock_double(flat.ctypes.data_as(C.POINTER(C.c_double)), a.ndim, dims)


def from_mx(p):
    lib = load()
    nd = lib.mock_ndim(p)
    shape = tuple(lib.mock_dim(p, i) for i in range(nd))
    n = int(np.prod(shape)) if shape else 0
    if n == 0:
        return np.zeros(shape)
    flat = np.ctypeslib.as_array(lib.mock_data(p), shape=(n,)).copy()
    return flat.reshape(shape, order="F")


def call(cmd, *args, nlhs=1):
    """[out1, ..] = gpz_b200_mex(cmd, args...)"""
    lib = load()
    prhs = [to_mx(cmd)] + [to_mx(a) for a in args]
    arr = (C.c_void_p * len(prhs))(*prhs)
    nout = max(nlhs, 1)
    plhs = (C.c_void_p * 8)()
    err = C.create_string_buffer(4096)
    rc = lib.mock_call(nlhs, plhs, len(prhs), arr, err, 4096)
    for p in prhs:
        lib.mock_free(p)
    if rc:
        ident, _, msg = err.value.decode(errors="replace").partition("|")
        raise MexError(ident, msg)
    outs = []
    for i in range(nout):
        if plhs[i]:
            outs.append(from_mx(plhs[i]))
            lib.mock_free(plhs[i])
        else:
            outs.append(None)
    return outs[0] if nlhs <= 1 else outs[:nlhs]


def printed(clear=True):
    lib = load()
    s = lib.mock_printed().decode()
    if clear:
        lib.mock_clear_printed()
    return s
