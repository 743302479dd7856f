"""ctypes driver of the MEX gateway built against the mini MEX runtime (csrc/mex_runtime): lets the tests call
mexFunction(nlhs, plhs, nrhs, prhs) the way MATLAB would, with numpy arrays standing in for MATLAB arrays
(logical MATLAB shape; data crosses in column-major order, complex interleaved as under -R2018a)."""
import ctypes as C

import numpy as np

import mamimo_b200 as mm


class MexError(RuntimeError):
    """what MATLAB would raise from mexErrMsgIdAndTxt: err.identifier / err.message"""

    def __init__(self, ident, msg):
        super().__init__("%s: %s" % (ident, msg))
        self.identifier, self.message = ident, msg


class Mex:
    def __init__(self):
        self.lib = C.CDLL(mm.build.build_mex_harness())
        L, vp = self.lib, C.c_void_p
        L.mexh_numeric.restype = vp
        L.mexh_numeric.argtypes = [C.c_int, C.POINTER(C.c_size_t), C.c_int, C.c_int]
        L.mexh_string.restype = vp
        L.mexh_string.argtypes = [C.c_char_p]
        L.mexh_struct.restype = vp
        L.mexh_set_field.argtypes = [vp, C.c_char_p, vp]
        L.mexh_data.restype = vp
        L.mexh_data.argtypes = [vp]
        L.mexh_ndim.argtypes = [vp]
        L.mexh_dim.restype = C.c_size_t
        L.mexh_dim.argtypes = [vp, C.c_int]
        L.mexh_is_single.argtypes = [vp]
        L.mexh_is_complex.argtypes = [vp]
        L.mexh_destroy.argtypes = [vp]
        L.mexh_call.argtypes = [C.c_int, C.POINTER(vp), C.c_int, C.POINTER(vp)]
        L.mexh_error_id.restype = C.c_char_p
        L.mexh_error_msg.restype = C.c_char_p

    # ---- numpy <-> mxArray
    def to_mx(self, v):
        L = self.lib
        if isinstance(v, str):
            return L.mexh_string(v.encode())
        if isinstance(v, dict):
            st = L.mexh_struct()
            for k, x in v.items():
                L.mexh_set_field(st, k.encode(), self.to_mx(x))
            return st
        a = np.asarray(v)
        if a.dtype not in (np.float32, np.float64, np.complex64, np.complex128):
            a = a.astype(np.float64)                      # MATLAB's default numeric class
        if a.ndim < 2:
            a = a.reshape((a.size, 1) if a.ndim == 1 else (1, 1))
        dims = (C.c_size_t * a.ndim)(*a.shape)
        mx = L.mexh_numeric(a.ndim, dims, int(a.dtype in (np.float32, np.complex64)), int(np.iscomplexobj(a)))
        flat = np.asfortranarray(a).ravel(order="F")
        if flat.size:
            C.memmove(L.mexh_data(mx), flat.ctypes.data, flat.nbytes)
        return mx

    def from_mx(self, mx):
        L = self.lib
        shape = tuple(int(L.mexh_dim(mx, i)) for i in range(L.mexh_ndim(mx)))
        single, cplx = bool(L.mexh_is_single(mx)), bool(L.mexh_is_complex(mx))
        dt = np.dtype({(False, False): np.float64, (True, False): np.float32, (False, True): np.complex128,
                       (True, True): np.complex64}[(single, cplx)])
        n = int(np.prod(shape))
        out = np.empty(n, dtype=dt)
        if n:
            C.memmove(out.ctypes.data, L.mexh_data(mx), out.nbytes)
        return out.reshape(shape, order="F")

    # ---- out1, out2, ... = mamimo_mex(cmd, args...)
    def call(self, *args, nlhs=0):
        L = self.lib
        prhs = [self.to_mx(a) for a in args]
        pin = (C.c_void_p * max(1, len(prhs)))(*prhs)
        pout = (C.c_void_p * max(1, nlhs))()
        try:
            rc = L.mexh_call(nlhs, pout, len(prhs), pin)
            if rc:
                raise MexError(L.mexh_error_id().decode(), L.mexh_error_msg().decode())
            outs = []
            for i in range(nlhs):
                if not pout[i]:
                    raise MexError("MATLAB:unassignedOutputs", "output %d was not assigned by the gateway" % (i + 1))
                outs.append(self.from_mx(pout[i]))
                L.mexh_destroy(pout[i])
            return outs[0] if nlhs == 1 else tuple(outs)
        finally:
            for p in prhs:
                L.mexh_destroy(p)

    def clear_mex(self):
        """`clear mex`: MATLAB runs the handlers registered with mexAtExit"""
        self.lib.mexh_clear_mex()

    def atexit_count(self):
        return int(self.lib.mexh_atexit_count())
