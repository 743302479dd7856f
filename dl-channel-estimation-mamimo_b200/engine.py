"""Host-side mirror of the reference's call surfaces over the C ABI (include/mamimo.h).

  Engine                      -- thin object wrapper of mamimo_engine (all math is in the CUDA library)
  helperMIMOChannelEstimate   -- MATLAB-shaped drop-in for pg/helperMIMOChannelEstimate.m:1 (LS part)
  CSIPredictor                -- drop-in for inference.py:6-68

numpy arrays are treated as HOST buffers (the library streams them through the GPU);
torch CUDA tensors are treated as DEVICE buffers (no copies, launched on torch's current stream).
"""
import ctypes as C
import os
import sys

import numpy as np

from . import _capi
from ._capi import lib, check


def _is_torch_cuda(x):
    return hasattr(x, "data_ptr") and getattr(x, "is_cuda", False)


def _np_ptr(a):
    return C.c_void_p(a.ctypes.data)


def pinned_empty(shape, dtype):
    """numpy array backed by pinned host memory from mamimo_host_alloc (freed with the array)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = lib.mamimo_host_alloc(max(n, 1))
    if not p:
        raise MemoryError("mamimo_host_alloc(%d) failed" % n)
    buf = (C.c_char * max(n, 1)).from_address(p)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    class _Owner:
        def __init__(self, ptr):
            self.ptr = ptr

        def __del__(self):
            lib.mamimo_host_free(self.ptr)

    # keep the allocation alive as long as any view of arr lives
    holder = _Owner(p)
    arr = arr.view(_PinnedArray)
    arr._holder = holder
    return arr


class _PinnedArray(np.ndarray):
    def __array_finalize__(self, obj):
        self._holder = getattr(obj, "_holder", None)


class Engine:
    """One engine per device.  See include/mamimo.h for the contract of every call."""

    def __init__(self, n_tx, n_rx, n_sc, n_ltf=None, n_ps=1, hidden=(1024, 1024), d_in=None, d_out=None,
                 input_mode="ls", precision="fp16x3", max_pkts=0, device=0, len_ltf=0, act_scale_log2=0,
                 mlp=True, kb_per_chunk=0, host_chunk_pkts=0, fc_single_cta=False, fc_sm_reserve=0):
        cfg = _capi.Config()
        lib.mamimo_config_init(C.byref(cfg))
        cfg.device = device
        cfg.n_tx, cfg.n_rx, cfg.n_sc = n_tx, n_rx, n_sc
        cfg.n_ltf = n_tx if n_ltf is None else n_ltf
        cfg.n_ps = n_ps
        cfg.input_mode = _capi.INPUT_MODES[input_mode]
        cfg.precision = _capi.PRECISIONS[precision]
        hidden = tuple(hidden) if mlp else ()
        if len(hidden) > _capi.MAX_HIDDEN:
            raise ValueError("at most %d hidden layers" % _capi.MAX_HIDDEN)
        cfg.n_hidden = len(hidden)
        for i, h in enumerate(hidden):
            cfg.hidden[i] = int(h)
        if mlp:
            cfg.d_in = int(d_in if d_in is not None else (n_sc if input_mode == "ls" else len_ltf + n_tx))
            cfg.d_out = int(d_out if d_out is not None else n_sc)
        cfg.len_ltf = len_ltf
        cfg.max_pkts = max_pkts
        cfg.act_scale_log2 = act_scale_log2
        cfg.kb_per_chunk = kb_per_chunk
        cfg.host_chunk_pkts = host_chunk_pkts
        cfg.fc_single_cta = 1 if fc_single_cta else 0
        cfg.fc_sm_reserve = fc_sm_reserve
        self.cfg = cfg
        self.precision = precision
        self.input_mode = input_mode
        self._h = C.c_void_p()
        check(lib.mamimo_create(C.byref(cfg), C.byref(self._h)))
        self.rows_per_pkt = n_tx * n_rx

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib.mamimo_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ------------------------------------------------------------------ tables / weights
    def set_pilots(self, x_pilot=None, P=None):
        """x_pilot [n_pilots] (ltf(ind), helperMIMOChannelEstimate.m:27) and P [n_tx, n_ltf] (helperGetP, :13)."""
        xp = None if x_pilot is None else np.ascontiguousarray(np.asarray(x_pilot, dtype=np.complex128))
        Pm = None if P is None else np.ascontiguousarray(np.asarray(P, dtype=np.complex128))
        if Pm is not None and Pm.shape != (self.cfg.n_tx, self.cfg.n_ltf):
            raise ValueError("P must be [n_tx, n_ltf]")
        n_pil = (self.cfg.n_sc + self.cfg.n_ps - 1) // self.cfg.n_ps
        if xp is not None and xp.shape != (n_pil,):
            raise ValueError("x_pilot must have %d entries" % n_pil)
        # double-precision tables: complex128 LS calls compute in FP64 end to end, everything else rounds them to FP32
        check(lib.mamimo_set_pilots_f64(self._h, None if xp is None else _np_ptr(xp), None if Pm is None else _np_ptr(Pm)),
              self._h)

    def load_weights(self, nets):
        """nets = {'real': layers, 'imag': layers}; layer = dict(W [in,out] (Keras kernel), b [out],
        bn = None | (gamma, beta, moving_mean, moving_var)).  Mirrors Model.load_weights (..._DNN.py:334)."""
        f32 = lambda t: np.ascontiguousarray(np.asarray(t, dtype=np.float32))
        for net, name in enumerate(("real", "imag")):
            for li, L in enumerate(nets[name]):
                W, b = f32(L["W"]), f32(L["b"])
                bn = L.get("bn")
                bnp = [f32(t) for t in bn] if bn is not None else [None] * 4
                args = [_np_ptr(W), _np_ptr(b)] + [None if t is None else _np_ptr(t) for t in bnp]
                check(lib.mamimo_load_layer(self._h, net, li, *args), self._h)
        check(lib.mamimo_finalize_weights(self._h), self._h)

    # ------------------------------------------------------------------ raw pointer calls (bench / C-ABI parity tests)
    def estimate_raw(self, y_ptr, y_type, n_pkt, hls_ptr, hr_ptr, hi_ptr, mem, stream=0):
        check(lib.mamimo_estimate(self._h, C.c_void_p(y_ptr), y_type, n_pkt, C.c_void_p(hls_ptr) if hls_ptr else None,
                                  C.c_void_p(hr_ptr), C.c_void_p(hi_ptr), mem, C.c_void_p(stream) if stream else None),
              self._h)

    STAGE_LS, STAGE_NET_REAL, STAGE_NET_IMAG, STAGE_GATHER = 1, 2, 4, 8

    # ------------------------------------------------------------------ fused all-gather (multi-GPU)
    def gather_create(self, world, rank, pkts_per_rank):
        """Allocate this rank's gathered planes; returns their device pointers (real, imag)."""
        pr, pi = C.c_void_p(), C.c_void_p()
        check(lib.mamimo_gather_create(self._h, world, rank, pkts_per_rank, C.byref(pr), C.byref(pi)), self._h)
        self._gather = (world, rank, pkts_per_rank, pr.value, pi.value)
        return pr.value, pi.value

    def gather_connect(self, real_ptrs, imag_ptrs):
        """Device pointers of EVERY rank's gathered planes as mapped in this process (entry [rank] = own)."""
        n = len(real_ptrs)
        ar = (C.c_void_p * n)(*real_ptrs)
        ai = (C.c_void_p * n)(*imag_ptrs)
        check(lib.mamimo_gather_connect(self._h, ar, ai), self._h)

    def gather_attach(self, world, rank, pkts_per_rank, real_ptrs, imag_ptrs, mc_real=0, mc_imag=0):
        """Gather over planes the caller allocated and mapped (mamimo_gather_attach); mc_* = NVSwitch multicast
        addresses of the two planes or 0."""
        ar = (C.c_void_p * world)(*real_ptrs)
        ai = (C.c_void_p * world)(*imag_ptrs)
        check(lib.mamimo_gather_attach(self._h, world, rank, pkts_per_rank, ar, ai, C.c_void_p(mc_real or None),
                                       C.c_void_p(mc_imag or None)), self._h)
        self._gather = (world, rank, pkts_per_rank, int(real_ptrs[rank]), int(imag_ptrs[rank]))

    def gather_planes(self):
        """This rank's gathered planes as torch CUDA tensors [world * pkts_per_rank * n_rx*n_tx, d_out] (no copy)."""
        import torch
        world, rank, ppr, pr, pi = self._gather
        shape = (world * ppr * self.rows_per_pkt, self.cfg.d_out)

        class _Raw:
            def __init__(self, ptr):
                self.__cuda_array_interface__ = {"shape": shape, "typestr": "<f4", "data": (ptr, False), "version": 2}

        dev = "cuda:%d" % self.cfg.device
        return torch.as_tensor(_Raw(pr), device=dev), torch.as_tensor(_Raw(pi), device=dev)

    def estimate_stages_raw(self, stages, y_ptr, y_type, n_pkt, hls_ptr, hr_ptr, hi_ptr, stream=0):
        """Piecewise run on device buffers (mamimo_estimate_stages): lets the caller overlap e.g. the all-gather
        of the real plane with the imaginary net."""
        vp = lambda p: C.c_void_p(p) if p else None
        check(lib.mamimo_estimate_stages(self._h, vp(y_ptr), y_type, n_pkt, vp(hls_ptr), vp(hr_ptr), vp(hi_ptr),
                                         _capi.MEM_DEVICE, vp(stream), stages), self._h)

    def ls_estimate_raw(self, y_ptr, y_type, n_pkt, h_ptr, h_type, mem, stream=0):
        check(lib.mamimo_ls_estimate(self._h, C.c_void_p(y_ptr), y_type, mem, n_pkt, C.c_void_p(h_ptr), h_type, mem,
                                     C.c_void_p(stream) if stream else None), self._h)

    # ------------------------------------------------------------------ array-level calls
    def _y_info(self, Y):
        c = self.cfg
        if tuple(Y.shape[1:]) != (c.n_rx, c.n_ltf, c.n_sc):
            raise ValueError("Y must be [n_pkt, n_rx=%d, n_ltf=%d, n_sc=%d], got %s"
                             % (c.n_rx, c.n_ltf, c.n_sc, tuple(Y.shape)))
        return int(Y.shape[0])

    def ls_estimate(self, Y):
        """Y [n_pkt, n_rx, n_ltf, n_sc] complex64/complex128 -> H_ls [n_pkt, n_rx, n_tx, n_sc] (same dtype)."""
        n_pkt = self._y_info(Y)
        c = self.cfg
        if _is_torch_cuda(Y):
            import torch
            if Y.dtype not in (torch.complex64, torch.complex128):
                raise TypeError("Y must be complex64 or complex128")
            Y = Y.contiguous()
            H = torch.empty((n_pkt, c.n_rx, c.n_tx, c.n_sc), dtype=Y.dtype, device=Y.device)
            t = _capi.C128 if Y.dtype == torch.complex128 else _capi.C64
            self.ls_estimate_raw(Y.data_ptr(), t, n_pkt, H.data_ptr(), t, _capi.MEM_DEVICE,
                                 torch.cuda.current_stream(Y.device).cuda_stream)
            return H
        Y = np.ascontiguousarray(Y)
        if Y.dtype not in (np.complex64, np.complex128):
            raise TypeError("Y must be complex64 or complex128")
        t = _capi.C128 if Y.dtype == np.complex128 else _capi.C64
        H = np.empty((n_pkt, c.n_rx, c.n_tx, c.n_sc), dtype=Y.dtype)
        self.ls_estimate_raw(Y.ctypes.data, t, n_pkt, H.ctypes.data, t, _capi.MEM_HOST)
        return H

    def poll_flags(self, stream=0):
        """Wait for `stream` and raise MamimoError if a kernel of an earlier DEVICE-buffer call latched a range /
        timeout / not-positive-definite condition (mamimo_poll_flags; device calls themselves only enqueue work)."""
        check(lib.mamimo_poll_flags(self._h, C.c_void_p(stream) if stream else None), self._h)

    def estimate(self, Y, want_ls=False, check_flags=True):
        """Full path (mode C).  Returns (H_real, H_imag[, H_ls]); H_* float32 [n_pkt*n_rx*n_tx, d_out].
        torch CUDA input: asynchronous on torch's current stream unless check_flags (default), which waits for the
        stream and raises on a latched device error instead of handing back garbage."""
        n_pkt = self._y_info(Y)
        c = self.cfg
        rows = n_pkt * self.rows_per_pkt
        if _is_torch_cuda(Y):
            import torch
            if Y.dtype not in (torch.complex64, torch.complex128):
                raise TypeError("Y must be complex64 or complex128")
            Y = Y.contiguous()
            t = _capi.C128 if Y.dtype == torch.complex128 else _capi.C64
            Hr = torch.empty((rows, c.d_out), dtype=torch.float32, device=Y.device)
            Hi = torch.empty_like(Hr)
            Hls = torch.empty((n_pkt, c.n_rx, c.n_tx, c.n_sc), dtype=torch.complex64, device=Y.device) if want_ls else None
            st = torch.cuda.current_stream(Y.device).cuda_stream
            self.estimate_raw(Y.data_ptr(), t, n_pkt, Hls.data_ptr() if want_ls else 0, Hr.data_ptr(), Hi.data_ptr(),
                              _capi.MEM_DEVICE, st)
            if check_flags:
                self.poll_flags(st)
        else:
            Y = np.ascontiguousarray(Y)
            if Y.dtype not in (np.complex64, np.complex128):
                raise TypeError("Y must be complex64 or complex128")
            t = _capi.C128 if Y.dtype == np.complex128 else _capi.C64
            Hr = np.empty((rows, c.d_out), dtype=np.float32)
            Hi = np.empty_like(Hr)
            Hls = np.empty((n_pkt, c.n_rx, c.n_tx, c.n_sc), dtype=np.complex64) if want_ls else None
            self.estimate_raw(Y.ctypes.data, t, n_pkt, Hls.ctypes.data if want_ls else 0, Hr.ctypes.data,
                              Hi.ctypes.data, _capi.MEM_HOST)
        return (Hr, Hi, Hls) if want_ls else (Hr, Hi)

    def predict_planes(self, X_real, X_imag, check_flags=True):
        """Mode B: float32 [rows, d_in] x2 -> float32 [rows, d_out] x2 (inference.py:29-30)."""
        c = self.cfg
        if _is_torch_cuda(X_real):
            import torch
            if not _is_torch_cuda(X_imag) or X_real.ndim != 2 or X_real.shape[1] != c.d_in or X_imag.shape != X_real.shape:
                raise ValueError("planes must be two CUDA tensors [rows, d_in=%d]" % c.d_in)
            Xr, Xi = X_real.contiguous().float(), X_imag.contiguous().float()
            rows = int(Xr.shape[0])
            Yr = torch.empty((rows, c.d_out), dtype=torch.float32, device=Xr.device)
            Yi = torch.empty_like(Yr)
            check(lib.mamimo_predict_planes(self._h, C.c_void_p(Xr.data_ptr()), C.c_void_p(Xi.data_ptr()), rows,
                                            C.c_void_p(Yr.data_ptr()), C.c_void_p(Yi.data_ptr()), _capi.MEM_DEVICE,
                                            C.c_void_p(torch.cuda.current_stream(Xr.device).cuda_stream)), self._h)
            if check_flags:
                self.poll_flags(torch.cuda.current_stream(Xr.device).cuda_stream)
            return Yr, Yi
        Xr = np.ascontiguousarray(X_real, dtype=np.float32)
        Xi = np.ascontiguousarray(X_imag, dtype=np.float32)
        if Xr.ndim != 2 or Xr.shape[1] != c.d_in or Xi.shape != Xr.shape:
            raise ValueError("planes must be [rows, d_in=%d]" % c.d_in)
        rows = Xr.shape[0]
        Yr = np.empty((rows, c.d_out), dtype=np.float32)
        Yi = np.empty_like(Yr)
        check(lib.mamimo_predict_planes(self._h, _np_ptr(Xr), _np_ptr(Xi), rows, _np_ptr(Yr), _np_ptr(Yi),
                                        _capi.MEM_HOST, None), self._h)
        return Yr, Yi

    def predict_time(self, sig_real, sig_imag):
        """Mode A: float32 [n_pkt, n_rx, len_ltf] x2 -> float32 [n_pkt*n_rx*n_tx, d_out] x2."""
        c = self.cfg
        Sr = np.ascontiguousarray(sig_real, dtype=np.float32)
        Si = np.ascontiguousarray(sig_imag, dtype=np.float32)
        if Sr.ndim != 3 or tuple(Sr.shape[1:]) != (c.n_rx, c.len_ltf) or Si.shape != Sr.shape:
            raise ValueError("signals must be [n_pkt, n_rx=%d, len_ltf=%d]" % (c.n_rx, c.len_ltf))
        n_pkt = Sr.shape[0]
        Yr = np.empty((n_pkt * self.rows_per_pkt, c.d_out), dtype=np.float32)
        Yi = np.empty_like(Yr)
        check(lib.mamimo_predict_time(self._h, _np_ptr(Sr), _np_ptr(Si), n_pkt, _np_ptr(Yr), _np_ptr(Yi),
                                      _capi.MEM_HOST, None), self._h)
        return Yr, Yi

    # ------------------------------------------------------------------ OFDM front-end
    def set_ofdm(self, fft_len, cp_len, sym_offset, carriers_1based):
        """ofdmdemod parameters (pg/generate_maMIMO_LTF.m:336-338); carriers = prm.CarriersLocations (1-based)."""
        car = np.ascontiguousarray(np.asarray(carriers_1based, dtype=np.int32).ravel())
        if car.size != self.cfg.n_sc:
            raise ValueError("need n_sc=%d carrier indices" % self.cfg.n_sc)
        check(lib.mamimo_set_ofdm(self._h, fft_len, cp_len, sym_offset, car.ctypes.data_as(C.POINTER(C.c_int32))), self._h)
        self._sym_len = fft_len + cp_len

    def _x_info(self, x):
        c = self.cfg
        want = (c.n_rx, c.n_ltf * self._sym_len)
        if x.ndim != 3 or tuple(x.shape[1:]) != want:
            raise ValueError("x must be [n_pkt, n_rx=%d, n_ltf*(fft+cp)=%d]" % want)
        if x.dtype not in (np.complex64, np.complex128):
            raise TypeError("x must be complex64 or complex128")
        return int(x.shape[0]), (_capi.C128 if x.dtype == np.complex128 else _capi.C64)

    def ofdm_demod(self, x):
        """time-domain x [n_pkt, n_rx, n_ltf*(fft+cp)] -> Y complex64 [n_pkt, n_rx, n_ltf, n_sc] (host arrays)."""
        x = np.ascontiguousarray(x)
        n_pkt, t = self._x_info(x)
        c = self.cfg
        Y = np.empty((n_pkt, c.n_rx, c.n_ltf, c.n_sc), dtype=np.complex64)
        check(lib.mamimo_ofdm_demod(self._h, _np_ptr(x), t, n_pkt, _np_ptr(Y), _capi.MEM_HOST, None), self._h)
        return Y

    def estimate_time(self, x, want_ls=False):
        """demod -> LS -> FC from time-domain samples.  Returns (H_real, H_imag[, H_ls]); on an engine built with
        mlp=False returns H_ls only."""
        x = np.ascontiguousarray(x)
        n_pkt, t = self._x_info(x)
        c = self.cfg
        mlp = c.d_out > 0
        Hls = np.empty((n_pkt, c.n_rx, c.n_tx, c.n_sc), dtype=np.complex64) if (want_ls or not mlp) else None
        Hr = np.empty((n_pkt * self.rows_per_pkt, c.d_out), dtype=np.float32) if mlp else None
        Hi = np.empty_like(Hr) if mlp else None
        check(lib.mamimo_estimate_time(self._h, _np_ptr(x), t, n_pkt, None if Hls is None else _np_ptr(Hls),
                                       None if Hr is None else _np_ptr(Hr), None if Hi is None else _np_ptr(Hi),
                                       _capi.MEM_HOST, None), self._h)
        if not mlp:
            return Hls
        return (Hr, Hi, Hls) if want_ls else (Hr, Hi)

    # ------------------------------------------------------------------ LMMSE smoother
    def lmmse(self, H_ls, tau_rms, snr_db, check_flags=True):
        """LMMSE_ce over a batch (pg/LMMSE_ce.m:23-39 as called at pg/helperMIMOChannelEstimate.m:37-39).
        H_ls [n_pkt, n_rx, n_tx, n_sc] complex64/128 (numpy = host, torch CUDA = device); tau_rms scalar or [n_pkt];
        snr_db scalar, [n_rx] or [n_pkt, n_rx] (SNR(i) in dB).  Returns H_mmse, same shape / dtype / residence."""
        c = self.cfg
        if tuple(H_ls.shape[1:]) != (c.n_rx, c.n_tx, c.n_sc):
            raise ValueError("H_ls must be [n_pkt, n_rx=%d, n_tx=%d, n_sc=%d]" % (c.n_rx, c.n_tx, c.n_sc))
        n_pkt = int(H_ls.shape[0])
        tau = np.ascontiguousarray(np.broadcast_to(np.asarray(tau_rms, dtype=np.float64), (n_pkt,)))
        snr = np.ascontiguousarray(np.broadcast_to(np.asarray(snr_db, dtype=np.float64), (n_pkt, c.n_rx)))
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        if _is_torch_cuda(H_ls):
            import torch
            if H_ls.dtype not in (torch.complex64, torch.complex128):
                raise TypeError("H_ls must be complex64 or complex128")
            H_ls = H_ls.contiguous()
            out = torch.empty_like(H_ls)
            t = _capi.C128 if H_ls.dtype == torch.complex128 else _capi.C64
            check(lib.mamimo_lmmse(self._h, C.c_void_p(H_ls.data_ptr()), t, n_pkt, dp(tau), dp(snr),
                                   C.c_void_p(out.data_ptr()), t, _capi.MEM_DEVICE,
                                   C.c_void_p(torch.cuda.current_stream(H_ls.device).cuda_stream)), self._h)
            if check_flags:       # e.g. Rpp not positive definite in FP64
                self.poll_flags(torch.cuda.current_stream(H_ls.device).cuda_stream)
            return out
        H_ls = np.ascontiguousarray(H_ls)
        if H_ls.dtype not in (np.complex64, np.complex128):
            raise TypeError("H_ls must be complex64 or complex128")
        t = _capi.C128 if H_ls.dtype == np.complex128 else _capi.C64
        out = np.empty_like(H_ls)
        check(lib.mamimo_lmmse(self._h, _np_ptr(H_ls), t, n_pkt, dp(tau), dp(snr), _np_ptr(out), t, _capi.MEM_HOST, None),
              self._h)
        return out

    # ------------------------------------------------------------------ per-subcarrier SVD
    def svd(self, H, want_vectors=True, check_flags=True, out_double=None):
        """Singular values and dominant right singular vectors of every per-tone [n_rx x n_tx] channel matrix
        (pg/omphybweights.m:174-176: H = Hin.'; [~,~,v] = svd(H)).  H [n_pkt, n_rx, n_tx, n_sc] complex64/128 (numpy =
        host, torch CUDA = device) -> sigma [n_pkt, n_rx, n_sc] (descending along axis 1) and V1 [n_pkt, n_rx, n_tx, n_sc]
        with V1[p, r, :, k] = v_r of tone k (unique up to a phase; V1 V1^H is the projector onto the row space).
        out_double: precision of sigma / V1 (None = that of H; True = float64 / complex128 whatever H is -- the
        arithmetic is FP64 either way)."""
        c = self.cfg
        if tuple(H.shape[1:]) != (c.n_rx, c.n_tx, c.n_sc):
            raise ValueError("H must be [n_pkt, n_rx=%d, n_tx=%d, n_sc=%d]" % (c.n_rx, c.n_tx, c.n_sc))
        n_pkt = int(H.shape[0])
        if _is_torch_cuda(H):
            import torch
            if H.dtype not in (torch.complex64, torch.complex128):
                raise TypeError("H must be complex64 or complex128")
            H = H.contiguous()
            dbl = H.dtype == torch.complex128
            odbl = dbl if out_double is None else bool(out_double)
            sig = torch.empty((n_pkt, c.n_rx, c.n_sc), dtype=torch.float64 if odbl else torch.float32, device=H.device)
            V = torch.empty(H.shape, dtype=torch.complex128 if odbl else torch.complex64, device=H.device) if want_vectors else None
            t = _capi.C128 if dbl else _capi.C64
            to = _capi.C128 if odbl else _capi.C64
            st = torch.cuda.current_stream(H.device).cuda_stream
            check(lib.mamimo_svd(self._h, C.c_void_p(H.data_ptr()), t, n_pkt, C.c_void_p(sig.data_ptr()),
                                 C.c_void_p(V.data_ptr()) if want_vectors else None, to, _capi.MEM_DEVICE, C.c_void_p(st)), self._h)
            if check_flags:
                self.poll_flags(st)
            return (sig, V) if want_vectors else sig
        H = np.ascontiguousarray(H)
        if H.dtype not in (np.complex64, np.complex128):
            raise TypeError("H must be complex64 or complex128")
        dbl = H.dtype == np.complex128
        odbl = dbl if out_double is None else bool(out_double)
        t = _capi.C128 if dbl else _capi.C64
        to = _capi.C128 if odbl else _capi.C64
        sig = np.empty((n_pkt, c.n_rx, c.n_sc), dtype=np.float64 if odbl else np.float32)
        V = np.empty(H.shape, dtype=np.complex128 if odbl else np.complex64) if want_vectors else None
        check(lib.mamimo_svd(self._h, _np_ptr(H), t, n_pkt, _np_ptr(sig), _np_ptr(V) if want_vectors else None, to,
                             _capi.MEM_HOST, None), self._h)
        return (sig, V) if want_vectors else sig

    # ------------------------------------------------------------------ OMP hybrid precoder (next row f-4, after svd)
    def set_steering_dictionary(self, At):
        """At [n_tx, n_rays] complex: the collection of candidate analog weights, column per ray, as MATLAB holds it
        (pg/BER_test_maMIMO_LTF.m:365 `At = steervec(prm.posTxElem,txang)`)."""
        At = np.asarray(At)
        if At.ndim != 2 or At.shape[0] != self.cfg.n_tx:
            raise ValueError("At must be [n_tx=%d, n_rays]" % self.cfg.n_tx)
        rows = np.ascontiguousarray(At.T.astype(np.complex128))            # [n_rays][n_tx]
        check(lib.mamimo_set_steering_dictionary(self._h, rows.view(np.float64).ctypes.data_as(C.POINTER(C.c_double)),
                                                 int(rows.shape[0])), self._h)
        self._n_rays = int(rows.shape[0])

    def omp(self, F, ns, n_rf, check_flags=True):
        """Orthogonal matching pursuit of every tone's Fopt over the steering dictionary (pg/omphybweights.m:178-179 +
        pg/ompdecomp.m:101-121).  F [n_pkt, f_rows >= ns, n_tx, n_sc] complex64/128 (rows 0..ns-1 = Fopt's columns; the V1
        of svd() as it stands), numpy = host / torch CUDA = device ->
        idx int32 [n_pkt, n_rf, n_sc] (0-based dictionary column, -1 after an early stop), err float32 [n_pkt, n_rf, n_sc]
        (residual norm after each round), Fbb complex [n_pkt, ns, n_rf, n_sc] (= the reference's Fbb(k, s, j))."""
        c = self.cfg
        if F.ndim != 4 or tuple(F.shape[2:]) != (c.n_tx, c.n_sc) or F.shape[1] < ns:
            raise ValueError("F must be [n_pkt, f_rows >= ns, n_tx=%d, n_sc=%d]" % (c.n_tx, c.n_sc))
        n_pkt, f_rows = int(F.shape[0]), int(F.shape[1])
        if _is_torch_cuda(F):
            import torch
            if F.dtype not in (torch.complex64, torch.complex128):
                raise TypeError("F must be complex64 or complex128")
            F = F.contiguous()
            t = _capi.C128 if F.dtype == torch.complex128 else _capi.C64
            idx = torch.empty((n_pkt, n_rf, c.n_sc), dtype=torch.int32, device=F.device)
            err = torch.empty((n_pkt, n_rf, c.n_sc), dtype=torch.float32, device=F.device)
            Fbb = torch.empty((n_pkt, ns, n_rf, c.n_sc), dtype=F.dtype, device=F.device)
            st = torch.cuda.current_stream(F.device).cuda_stream
            check(lib.mamimo_omp(self._h, C.c_void_p(F.data_ptr()), t, f_rows, n_pkt, ns, n_rf, C.c_void_p(idx.data_ptr()),
                                 C.c_void_p(err.data_ptr()), C.c_void_p(Fbb.data_ptr()), t, _capi.MEM_DEVICE, C.c_void_p(st)),
                  self._h)
            if check_flags:
                self.poll_flags(st)
            return idx, err, Fbb
        F = np.ascontiguousarray(F)
        if F.dtype not in (np.complex64, np.complex128):
            raise TypeError("F must be complex64 or complex128")
        t = _capi.C128 if F.dtype == np.complex128 else _capi.C64
        idx = np.empty((n_pkt, n_rf, c.n_sc), dtype=np.int32)
        err = np.empty((n_pkt, n_rf, c.n_sc), dtype=np.float32)
        Fbb = np.empty((n_pkt, ns, n_rf, c.n_sc), dtype=F.dtype)
        check(lib.mamimo_omp(self._h, _np_ptr(F), t, f_rows, n_pkt, ns, n_rf, _np_ptr(idx), _np_ptr(err), _np_ptr(Fbb), t,
                             _capi.MEM_HOST, None), self._h)
        return idx, err, Fbb

    def omp_precoder(self, H, ns, n_rf):
        """[Fbb, Frf] = omphybweights(hD, Ns, NtRF, AtExp) for a batch (pg/BER_test_maMIMO_LTF.m:372): svd() then omp().
        H [n_pkt, n_rx, n_tx, n_sc] -> (idx, err, Fbb) as omp(); the reference's Frf(k, j, :) is dictionary column
        idx[p, j, k]."""
        _, V = self.svd(H, out_double=True)        # complex128 vectors: the column choice must not ride on FP32 rounding
        idx, err, Fbb = self.omp(V, ns, n_rf)
        return idx, err, Fbb

    def debug_counters(self, reset=True):
        """role counters of the CTA-pair FC kernel (mamimo_get_debug_counters; debug build + MAMIMO_FC_DEBUG=1)"""
        out = (C.c_uint64 * 8)()
        check(lib.mamimo_get_debug_counters(self._h, out, 1 if reset else 0), self._h)
        return [int(v) for v in out]

    def synchronize(self):
        check(lib.mamimo_synchronize(self._h), self._h)

    def profile_begin(self):
        check(lib.mamimo_profile_begin(self._h), self._h)

    def profile_end(self):
        p = _capi.Profile()
        check(lib.mamimo_profile_end(self._h, C.byref(p)), self._h)
        return dict(ls_ms=p.ls_ms, fc_ms=p.fc_ms, stage_ms=p.stage_ms, ls_launches=int(p.ls_launches),
                    fc_launches=int(p.fc_launches), stage_launches=int(p.stage_launches),
                    lmmse_ms=p.lmmse_ms, lmmse_launches=int(p.lmmse_launches))

    def stats(self):
        s = _capi.Stats()
        check(lib.mamimo_get_stats(self._h, C.byref(s)), self._h)
        return dict(kernel_launches=int(s.kernel_launches), h2d_bytes=int(s.h2d_bytes), d2h_bytes=int(s.d2h_bytes),
                    last_device_flags=int(s.last_device_flags), graph_launches=int(s.graph_launches))


def ipc_export(dev_ptr):
    """64-byte CUDA IPC handle of a cudaMalloc'ed device pointer (to send to another process)."""
    buf = C.create_string_buffer(64)
    st = lib.mamimo_ipc_export(C.c_void_p(dev_ptr), buf)
    if st != _capi.OK:
        raise _capi.MamimoError(st, "cudaIpcGetMemHandle failed")
    return buf.raw


def ipc_open(handle):
    """Map another process's exported device allocation into this process; returns the device pointer."""
    p = C.c_void_p()
    st = lib.mamimo_ipc_open(C.c_char_p(handle), C.byref(p))
    if st != _capi.OK:
        raise _capi.MamimoError(st, "cudaIpcOpenMemHandle failed")
    return p.value


# ---------------------------------------------------------------------- integer tables
def vht_ltf256():
    out = (C.c_int8 * 256)()
    lib.mamimo_vht_ltf256(out)
    return np.frombuffer(out, dtype=np.int8).copy()


def carriers_locations():
    n = lib.mamimo_carriers_locations(None, 0)
    out = (C.c_int32 * n)()
    lib.mamimo_carriers_locations(out, n)
    return np.frombuffer(out, dtype=np.int32).copy()


def default_p(n):
    out = np.empty((n, n), dtype=np.float32)
    st = lib.mamimo_default_p(n, out.ctypes.data_as(C.POINTER(C.c_float)))
    if st != _capi.OK:
        raise ValueError("default P needs a power-of-two size")
    return out


def pair_row(p, i_rx, i_tx, n_rx, n_tx):
    return int(lib.mamimo_pair_row(p, i_rx, i_tx, n_rx, n_tx))


def tau_rms(h):
    """rms delay LMMSE_ce.m:27-30 derives from its `h` argument (real or complex vector)."""
    h = np.asarray(h).ravel()
    cplx = np.iscomplexobj(h)
    buf = np.ascontiguousarray(h.astype(np.complex128).view(np.float64) if cplx else h.astype(np.float64))
    return float(lib.mamimo_tau_rms(buf.ctypes.data_as(C.POINTER(C.c_double)), h.size, int(cplx)))


# ---------------------------------------------------------------------- MATLAB-shaped drop-in
_ENGINE_CACHE = {}


def helperMIMOChannelEstimate(rxData, prm, Nps=1, tau=None, SNR=None, isMMSE=False, P=None, ltf=None, device=0):
    """[hD, P, ltf_o, hDmmse] = helperMIMOChannelEstimate(rxData, prm, Nps, tau, SNR, isMMSE)

    Same argument meaning as pg/helperMIMOChannelEstimate.m:1.  rxData complex [Nsc, nltf, Nr]
    (MATLAB logical shape) or [Nsc, nltf, Nr, Npkt] for a batch hoisted out of the packet loop;
    prm needs 'numSTS' and 'CarriersLocations' (1-based, :9,26).  Returns hD [Nsc, numSTS, Nr(, Npkt)]
    complex128 -- computed in FP64 on the device like MATLAB's (rxData is handed over as complex128 and the engine's
    complex128-in / complex128-out LS runs in double) --, P, ltf_o = ltf(ind) (:29) and hDmmse (zeros unless isMMSE,
    as in the reference :32).  With
    isMMSE, tau is LMMSE_ce's `h` vector (one per call, or a list of Npkt vectors for a batch) and SNR is SNR(i)
    in dB, [Nr] or [Nr, Npkt] (:37-39).
    """
    if isMMSE and (tau is None or SNR is None):
        raise ValueError("isMMSE needs tau and SNR (helperMIMOChannelEstimate.m:38)")
    get = (lambda k: prm[k]) if isinstance(prm, dict) else (lambda k: getattr(prm, k))
    num_sts = int(get("numSTS"))
    ind = np.asarray(get("CarriersLocations"), dtype=np.int64).ravel()
    rx = np.asarray(rxData)
    squeeze = rx.ndim == 3
    if squeeze:
        rx = rx[..., None]
    if rx.ndim != 4:
        raise ValueError("rxData must be [Nsc, nltf, Nr] or [Nsc, nltf, Nr, Npkt]")
    nsc, nltf, nrx, npkt = rx.shape
    if nsc != ind.size:
        raise ValueError("size(rxData,1) must equal numel(prm.CarriersLocations)")
    if nltf != num_sts:
        raise ValueError("nltf should be == numSTS (helperMIMOChannelEstimate.m:10)")
    table = vht_ltf256() if ltf is None else np.asarray(ltf)
    ltf_o = table[ind - 1].astype(np.float64)
    Pm = default_p(num_sts).astype(np.float64) if P is None else np.asarray(P)[:num_sts, :num_sts]
    key = (num_sts, nrx, nsc, device)
    eng = _ENGINE_CACHE.get(key)
    if eng is None:
        eng = _ENGINE_CACHE[key] = Engine(num_sts, nrx, nsc, mlp=False, device=device)
    eng.set_pilots(ltf_o, Pm)
    # MATLAB [Nsc, nltf, Nr, Npkt] column-major == C-order [Npkt, Nr, nltf, Nsc]
    Y = np.ascontiguousarray(np.transpose(rx, (3, 2, 1, 0)).astype(np.complex128, copy=False))
    H = eng.ls_estimate(Y)                                   # [Npkt, Nr, Nt, Nsc]
    hD = np.transpose(H, (3, 2, 1, 0))
    if isMMSE:
        # the reference applies Nps inside LMMSE_ce only (:38; LS itself always sees every tone): a second engine
        # carries the pilot spacing of the smoother
        if int(Nps) != 1:
            lkey = (num_sts, nrx, nsc, device, "lmmse", int(Nps))
            leng = _ENGINE_CACHE.get(lkey)
            if leng is None:
                leng = _ENGINE_CACHE[lkey] = Engine(num_sts, nrx, nsc, n_ps=int(Nps), mlp=False, device=device)
        else:
            leng = eng
        taus = tau if (isinstance(tau, (list, tuple)) and len(tau) == npkt and np.ndim(tau[0]) > 0) else [tau] * npkt
        t_rms = np.array([tau_rms(t) for t in taus])
        snr = np.asarray(SNR, dtype=np.float64)
        snr = np.broadcast_to(snr.reshape(nrx, -1).T, (npkt, nrx))     # [Nr] or [Nr, Npkt] -> [Npkt, Nr]
        hM = np.transpose(leng.lmmse(H, t_rms, snr), (3, 2, 1, 0))
    else:
        hM = np.zeros_like(hD)
    if squeeze:
        hD, hM = hD[..., 0], hM[..., 0]
    return hD, Pm, ltf_o.reshape(-1, 1), hM


# ---------------------------------------------------------------------- inference.py drop-in
class CSIPredictor:
    """Drop-in for inference.py:6-68 backed by the CUDA engine (mode B).

    model_path holds 'real_weights.npz' / 'imag_weights.npz' (flat exports of the Keras layers:
    W0,b0[,bn0_gamma,bn0_beta,bn0_mean,bn0_var],W1,...; written by tools/keras_to_npz.py, TensorFlow/h5py are not
    needed to load them), or the reference's own hdf5 / SavedModel artefacts when h5py / tensorflow are installed,
    or the nets are given directly as nets={'real': layers, 'imag': layers}.
    """

    def __init__(self, model_path=None, experiment="RICE_RENEW", verbose=False, nets=None, precision="fp16x3",
                 device=0):
        self.path = model_path
        self.experiment = experiment
        self.verbose = verbose
        self.nets = nets if nets is not None else self.load_model()
        real = self.nets["real"]
        d_in = int(np.asarray(real[0]["W"]).shape[0])
        hidden = [int(np.asarray(L["W"]).shape[1]) for L in real[:-1]]
        d_out = int(np.asarray(real[-1]["W"]).shape[1])
        self.engine = Engine(1, 1, 1, n_ltf=1, hidden=hidden, d_in=d_in, d_out=d_out, input_mode="planes",
                             precision=precision, device=device)
        self.engine.load_weights(self.nets)
        if verbose:
            for name in ("real", "imag"):
                print("------- %s Model Summary -------" % name.capitalize())
                for i, L in enumerate(self.nets[name]):
                    print("  dense%d %s bn=%s" % (i, np.asarray(L["W"]).shape, L.get("bn") is not None))

    def load_model(self):
        """inference.py:14-22 loads <path>/{real,imag}_keras_model; here, in order: <d>_weights.npz (tools/keras_to_npz.py),
        <d>_weights-improvement.hdf5 (needs h5py), <d>_keras_model (needs tensorflow) -- see weights.py."""
        from . import weights
        return weights.load_nets(self.path)

    def inference(self, input_batch):
        X = self.preprocess_data(input_batch)
        out_r, out_i = self.engine.predict_planes(X.real, X.imag)        # inference.py:29-30
        return self.postprocess_data(out_r + 1j * out_i)                  # :31-32

    def preprocess_data(self, input_batch):
        if self.experiment == "RICE_RENEW":
            if input_batch.dtype != np.complex128:                        # inference.py:41-43
                print("[CSIPredictor] ERROR: Input batch must be of type np.complex128")
                sys.exit(-1)
        return input_batch

    def postprocess_data(self, output_batch):
        postp = output_batch
        if self.experiment == "RICE_RENEW":
            if output_batch.shape[1] == 52:                               # inference.py:56-63
                n = output_batch.shape[0]
                tmp = np.concatenate((np.zeros((n, 6)), output_batch[:, 0:26], np.zeros((n, 1)),
                                      output_batch[:, 26:], np.zeros((n, 5))), axis=1)
                postp = np.fft.ifftshift(tmp, axes=1)
            else:                                                          # :64-66
                print("[CSIPredictor] ERROR: Output samples must have size 52 (assuming FFTLen = 64).")
                sys.exit(-1)
        return postp
