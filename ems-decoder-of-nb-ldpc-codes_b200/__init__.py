"""ems-decoder-of-nb-ldpc-codes_b200 -- B200-native EMS decoder for non-binary LDPC codes.

The product is ``libnbldpc_b200.so`` (plain-C host layer + hand-written sm_100a CUDA kernels behind the
C ABI declared in ``include/nbldpc_b200.h``).  This Python module is only a thin ctypes binding of that
ABI for the tests and ``bench.py``; it contains no decoding logic and no CPU fallback: when the shared
library is missing it raises, and every compute entry point fails with NBGPU_ECUDA without a B200.

Import with ``importlib.import_module("ems-decoder-of-nb-ldpc-codes_b200")`` (the directory name is the
repository's, hyphens included) or through ``nbldpc.py`` at the repo root.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NBLDPC_B200_LIB") or os.path.join(HERE, "libnbldpc_b200.so")   # override: tuning builds only

OK, EINVAL, EIO, ENOMEM, ECUDA, ESTATE, ERANK = 0, -1, -2, -3, -4, -5, -6
ALIST_AUTO, ALIST_UBS, ALIST_KN, ALIST_FULL = 0, 1, 2, 3

_ip = C.POINTER(C.c_int)
_fp = C.POINTER(C.c_float)
_lp = C.POINTER(C.c_long)


class NbgpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("nbgpu error %d: %s" % (code, msg))
        self.code = code


class Params(C.Structure):
    """nbgpu_params (include/nbldpc_b200.h)."""
    _fields_ = [("n_m", C.c_int), ("nb_oper", C.c_int), ("nb_iter_max", C.c_int), ("offset", C.c_float),
                ("ecn_kind", C.c_int), ("early_stop", C.c_int),
                ("d1", C.c_int), ("d2", C.c_int), ("d3", C.c_int), ("cfg_trunc", C.c_int), ("n_cv", C.c_int),
                ("border", C.c_int), ("frames_per_cta", C.c_int), ("cns_per_step", C.c_int)]


class Rng(C.Structure):
    _fields_ = [("x", C.c_uint64)]


def build(force=False):
    """Compile libnbldpc_b200.so in-tree (gcc + nvcc, sm_100a only)."""
    args = ["make", "-s", "-C", os.path.join(HERE, "csrc")]
    if force:
        args.append("-B")
    subprocess.run(args, check=True)


_lib = None


def apsk64_table():
    """The normalised 64-APSK constellation of ModelChannel_AWGN_64 (channel.c:133-222), [64][2], by binary image."""
    mod = np.zeros((64, 2), np.float32)
    _check(lib().nbgpu_apsk64_table(mod.ctypes.data_as(_fp)))
    return mod


def sigma_apsk64(ebn):
    return float(lib().nbgpu_sigma_apsk64(C.c_float(ebn)))


def lib():
    """The loaded C-ABI library.  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    sig = {
        "nbgpu_code_load": (C.c_int, [C.POINTER(vp), C.c_char_p, C.c_int]),
        "nbgpu_code_from_arrays": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int, C.c_int, _ip, _ip, _ip, _ip, _ip, _ip, _ip]),
        "nbgpu_code_free": (None, [vp]),
        "nbgpu_code_info": (None, [vp, _ip]),
        "nbgpu_code_rate": (C.c_float, [vp]),
        "nbgpu_code_graph": (None, [vp, _ip, _ip, _ip]),
        "nbgpu_code_tables": (None, [vp, _ip, _ip, _ip, _ip]),
        "nbgpu_rng_reference_default": (None, [C.POINTER(Rng)]),
        "nbgpu_rng_skip": (None, [C.POINTER(Rng), C.c_uint64]),
        "nbgpu_rng_drand48": (C.c_double, [C.POINTER(Rng)]),
        "nbgpu_code_prepare_encoder": (C.c_int, [vp]),
        "nbgpu_random_codeword": (C.c_int, [vp, C.POINTER(Rng), _ip, _ip]),
        "nbgpu_sigma": (C.c_float, [vp, C.c_float]),
        "nbgpu_awgn_bpsk_noise": (C.c_int, [vp, C.POINTER(Rng), _ip, C.c_float, _fp]),
        "nbgpu_apsk64_table": (C.c_int, [_fp]),
        "nbgpu_sigma_apsk64": (C.c_float, [C.c_float]),
        "nbgpu_awgn_apsk64_noise": (C.c_int, [vp, C.POINTER(Rng), _ip, C.c_float, _fp]),
        "nbgpu_channel_awgn_apsk64": (C.c_int, [vp, _fp, C.c_float, C.c_int, _fp, _fp, _ip]),
        "nbgpu_decode_apsk64": (C.c_int, [vp, _fp, C.c_float, C.c_int, _ip, _ip, _ip]),
        "nbgpu_upload_apsk64": (C.c_int, [vp, _fp, C.c_float, C.c_int]),
        "nbgpu_create": (C.c_int, [C.POINTER(vp), vp, C.POINTER(Params), C.c_int, C.c_int]),
        "nbgpu_destroy": (None, [vp]),
        "nbgpu_last_error": (C.c_char_p, [vp]),
        "nbgpu_channel_awgn_bpsk": (C.c_int, [vp, _fp, C.c_float, C.c_int, _fp, _fp, _ip]),
        "nbgpu_decode_noisy": (C.c_int, [vp, _fp, C.c_float, C.c_int, _ip, _ip, _ip]),
        "nbgpu_decode_llr": (C.c_int, [vp, _fp, C.c_int, _ip, _ip, _ip]),
        "nbgpu_upload_noisy": (C.c_int, [vp, _fp, C.c_float, C.c_int]),
        "nbgpu_upload_llr": (C.c_int, [vp, _fp, C.c_int]),
        "nbgpu_run": (C.c_int, [vp]),
        "nbgpu_sync": (C.c_int, [vp]),
        "nbgpu_download": (C.c_int, [vp, _ip, _ip, _ip]),
        "nbgpu_last_kernel_ms": (C.c_int, [vp, _fp]),
        "nbgpu_launch_count": (C.c_long, [vp]),
        "nbgpu_timer_begin": (C.c_int, [vp]),
        "nbgpu_timer_end": (C.c_int, [vp, _fp]),
        "nbgpu_host_register": (C.c_int, [vp, C.c_size_t]),
        "nbgpu_host_unregister": (C.c_int, [vp]),
        "nbgpu_geometry": (C.c_int, [vp, _ip]),
        "nbgpu_slow_selects": (C.c_long, [vp]),
        "nbgpu_get_state": (C.c_int, [vp, C.c_int, _fp, _fp]),
        "nbgpu_source_frames": (C.c_int, [vp, vp, C.POINTER(Rng), C.c_uint64, C.c_int, C.c_float]),
        "nbgpu_source_download": (C.c_int, [vp, _ip, _fp]),
        "nbgpu_source_results": (C.c_int, [vp, _ip, _ip, _ip]),
        "nbgpu_source_fixups": (C.c_long, [vp]),
        "nbgpu_source_set_margin": (C.c_int, [vp, C.c_double]),
        "nbgpu_check_node": (C.c_int, [vp, C.c_int, _fp, _ip, _fp, _ip, C.c_int]),
        "nbgpu_elementary_step": (C.c_int, [vp, _fp, _fp, _ip, _ip, _fp, _ip, C.c_int]),
        "nbgpu_select_nm": (C.c_int, [vp, _fp, _fp, _ip, C.c_int]),
        "nbgpu_decision_syndrome": (C.c_int, [vp, _fp, _ip, _ip, C.c_int]),
        "nbgpu_accumulate_stats": (C.c_int, [vp, _ip, _ip, _ip, _ip, C.c_int, _lp]),
        "nbgpu_accumulate_results": (C.c_int, [_ip, _ip, _ip, C.c_int, _lp]),
        "nbgpu_config_table": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _ip, C.c_int]),
        "nbgpu_version": (C.c_char_p, []),
        "nbgpu_device_count": (C.c_int, []),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


def _i(a):
    return a.ctypes.data_as(_ip) if a is not None else None


def _f(a):
    return a.ctypes.data_as(_fp) if a is not None else None


def _check(rc, ctx=None):
    if rc != OK:
        msg = lib().nbgpu_last_error(ctx)
        raise NbgpuError(rc, msg.decode(errors="replace") if msg else "")


class Code:
    """nbgpu_code: alist matrix + GF(q) tables (reference: code_t + table_t, LoadCode/LoadTables)."""

    def __init__(self, path=None, dialect=ALIST_AUTO, arrays=None):
        L = lib()
        self.h = C.c_void_p()
        if arrays is not None:
            a = arrays
            keep = [np.ascontiguousarray(a[k], np.int32) if a.get(k) is not None else None
                    for k in ("row_deg", "col", "val", "bingf", "addgf", "mulgf", "divgf")]
            _check(L.nbgpu_code_from_arrays(C.byref(self.h), int(a["N"]), int(a["M"]), int(a["q"]),
                                            *[_i(x) for x in keep]))
        else:
            _check(L.nbgpu_code_load(C.byref(self.h), os.fsencode(path), dialect))
        info = np.zeros(10, np.int32)
        L.nbgpu_code_info(self.h, _i(info))
        (self.N, self.M, self.K, self.q, self.logq, self.E, self.dc_max, self.dc_min, self.dialect) = [int(x) for x in info[:9]]
        self.rate = float(L.nbgpu_code_rate(self.h))
        self.row_deg = np.zeros(self.M, np.int32)
        self.col = np.zeros(self.E, np.int32)
        self.val = np.zeros(self.E, np.int32)
        L.nbgpu_code_graph(self.h, _i(self.row_deg), _i(self.col), _i(self.val))
        self.row_ptr = np.concatenate([[0], np.cumsum(self.row_deg)]).astype(np.int64)
        self.rng = Rng()
        L.nbgpu_rng_reference_default(C.byref(self.rng))

    def tables(self):
        q, lg = self.q, self.logq
        b = np.zeros((q, lg), np.int32)
        a = np.zeros((q, q), np.int32); m = np.zeros((q, q), np.int32); d = np.zeros((q, q), np.int32)
        lib().nbgpu_code_tables(self.h, _i(b), _i(a), _i(m), _i(d))
        return b, a, m, d

    @property
    def info_bits(self):
        return self.K * self.logq

    # frame source (host side of the reference: tools.c:124-268, channel.c:51-62)
    def rng_default(self):
        lib().nbgpu_rng_reference_default(C.byref(self.rng))

    def rng_skip(self, n):
        lib().nbgpu_rng_skip(C.byref(self.rng), int(n))

    def drand48(self):
        return lib().nbgpu_rng_drand48(C.byref(self.rng))

    def prepare_encoder(self):
        _check(lib().nbgpu_code_prepare_encoder(self.h))

    def random_codeword(self):
        cw = np.zeros(self.N, np.int32)
        nbin = np.zeros((self.N, self.logq), np.int32)
        _check(lib().nbgpu_random_codeword(self.h, C.byref(self.rng), _i(cw), _i(nbin)))
        return cw, nbin

    def sigma(self, ebn):
        return float(lib().nbgpu_sigma(self.h, C.c_float(ebn)))

    def noise(self, nbin, ebn):
        """channel.c:52-62 for one frame; nbin None = all-zero codeword."""
        noisy = np.zeros((self.N, self.logq), np.float32)
        nb = np.ascontiguousarray(nbin, np.int32) if nbin is not None else None
        _check(lib().nbgpu_awgn_bpsk_noise(self.h, C.byref(self.rng), _i(nb), C.c_float(ebn), _f(noisy)))
        return noisy

    def noise_apsk64(self, nbin, ebn):
        """channel.c:234-263 for one frame of a GF(64) code sent as 64-APSK symbols: noisy[N][2]."""
        noisy = np.zeros((self.N, 2), np.float32)
        nb = np.ascontiguousarray(nbin, np.int32) if nbin is not None else None
        _check(lib().nbgpu_awgn_apsk64_noise(self.h, C.byref(self.rng), _i(nb), C.c_float(ebn), _f(noisy)))
        return noisy

    def accumulate_stats(self, codeword_bits, decide, synd, iters, stats):
        cb = np.ascontiguousarray(codeword_bits, np.int32)
        d = np.ascontiguousarray(decide, np.int32)
        s = np.ascontiguousarray(synd, np.int32)
        it = np.ascontiguousarray(iters, np.int32)
        B = d.shape[0]
        assert stats.dtype == np.int64 and stats.shape == (6,)
        _check(lib().nbgpu_accumulate_stats(self.h, _i(cb), _i(d), _i(s), _i(it), B, stats.ctypes.data_as(_lp)))
        return stats

    def close(self):
        if self.h:
            lib().nbgpu_code_free(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Decoder:
    """nbgpu_ctx: the decode path NB_LDPC.c:266-474 for batches of frames on one B200."""

    def __init__(self, code, n_m, nb_oper, nb_iter_max, offset, early_stop=True, ecn_kind=0, device=0,
                 max_batch=1, d1=0, d2=0, d3=0, cfg_trunc=0, n_cv=0, border=0, frames_per_cta=0, cns_per_step=0):
        self.code = code
        self.p = Params(n_m, nb_oper, nb_iter_max, offset, ecn_kind, int(early_stop), d1, d2, d3, cfg_trunc, n_cv,
                        border, frames_per_cta, cns_per_step)
        self.h = C.c_void_p()
        self.max_batch = max_batch
        _check(lib().nbgpu_create(C.byref(self.h), code.h, C.byref(self.p), device, max_batch))
        self.B = 0

    def _out(self, B):
        return (np.zeros((B, self.code.N), np.int32), np.zeros(B, np.int32), np.zeros(B, np.int32))

    def decode_noisy(self, noisy, sigma):
        noisy = np.ascontiguousarray(noisy, np.float32).reshape(-1, self.code.N, self.code.logq)
        B = noisy.shape[0]
        d, s, it = self._out(B)
        _check(lib().nbgpu_decode_noisy(self.h, _f(noisy), C.c_float(sigma), B, _i(d), _i(s), _i(it)), self.h)
        self.B = B
        return d, s, it

    def decode_apsk64(self, noisy, sigma):
        """nbgpu_decode_apsk64: frames of a GF(64) code received as 64-APSK samples [B, N, 2]."""
        noisy = np.ascontiguousarray(noisy, np.float32).reshape(-1, self.code.N, 2)
        B = noisy.shape[0]
        d, s, it = self._out(B)
        _check(lib().nbgpu_decode_apsk64(self.h, _f(noisy), C.c_float(sigma), B, _i(d), _i(s), _i(it)), self.h)
        self.B = B
        return d, s, it

    def channel_apsk64(self, noisy, sigma, want_sorted=False):
        noisy = np.ascontiguousarray(noisy, np.float32).reshape(-1, self.code.N, 2)
        B = noisy.shape[0]
        llr = np.zeros((B, self.code.N, self.code.q), np.float32)
        il = np.zeros_like(llr) if want_sorted else None
        ig = np.zeros(llr.shape, np.int32) if want_sorted else None
        _check(lib().nbgpu_channel_awgn_apsk64(self.h, _f(noisy), C.c_float(sigma), B, _f(llr), _f(il), _i(ig)), self.h)
        return (llr, il, ig) if want_sorted else llr

    def decode_noisy_into(self, noisy, sigma, out):
        """nbgpu_decode_noisy on caller-owned (e.g. pinned) buffers, no allocation: noisy [B,N,logq] f32, out = (decide, synd, iters)."""
        B = noisy.shape[0]
        _check(lib().nbgpu_decode_noisy(self.h, _f(noisy), C.c_float(sigma), B, _i(out[0]), _i(out[1]), _i(out[2])), self.h)
        self.B = B
        return out

    def decode_llr(self, llr):
        llr = np.ascontiguousarray(llr, np.float32).reshape(-1, self.code.N, self.code.q)
        B = llr.shape[0]
        d, s, it = self._out(B)
        _check(lib().nbgpu_decode_llr(self.h, _f(llr), B, _i(d), _i(s), _i(it)), self.h)
        self.B = B
        return d, s, it

    def upload_noisy(self, noisy, sigma):
        noisy = np.ascontiguousarray(noisy, np.float32).reshape(-1, self.code.N, self.code.logq)
        self.B = noisy.shape[0]
        _check(lib().nbgpu_upload_noisy(self.h, _f(noisy), C.c_float(sigma), self.B), self.h)

    def upload_llr(self, llr):
        llr = np.ascontiguousarray(llr, np.float32).reshape(-1, self.code.N, self.code.q)
        self.B = llr.shape[0]
        _check(lib().nbgpu_upload_llr(self.h, _f(llr), self.B), self.h)

    def run(self):
        _check(lib().nbgpu_run(self.h), self.h)

    def sync(self):
        _check(lib().nbgpu_sync(self.h), self.h)

    def download(self, out=None):
        d, s, it = out if out is not None else self._out(self.B)
        _check(lib().nbgpu_download(self.h, _i(d), _i(s), _i(it)), self.h)
        return d, s, it

    def source_frames(self, frame0, B, ebn, origin=None):
        """nbgpu_source_frames: frames [frame0, frame0+B) of the reference's stream generated on the device and left resident
        (origin: drand48 state before frame 0 as an int, default = the reference's unseeded stream)."""
        r = Rng()
        lib().nbgpu_rng_reference_default(C.byref(r))
        if origin is not None:
            r.x = origin
        _check(lib().nbgpu_source_frames(self.h, self.code.h, C.byref(r), frame0, B, C.c_float(ebn)), self.h)
        self.B = B

    def source_download(self, want_noisy=True):
        cw = np.zeros((self.B, self.code.N), np.int32)
        noisy = np.zeros((self.B, self.code.N, self.code.logq), np.float32) if want_noisy else None
        _check(lib().nbgpu_source_download(self.h, _i(cw), _f(noisy)), self.h)
        return cw, noisy

    def source_results(self):
        """(bit_errors[B], synd[B], iters[B]) of the generated batch after run()"""
        e, s, it = (np.zeros(self.B, np.int32) for _ in range(3))
        _check(lib().nbgpu_source_results(self.h, _i(e), _i(s), _i(it)), self.h)
        return e, s, it

    def source_fixups(self):
        return int(lib().nbgpu_source_fixups(self.h))

    def source_set_margin(self, margin):
        _check(lib().nbgpu_source_set_margin(self.h, C.c_double(margin)), self.h)

    def last_kernel_ms(self):
        ms = C.c_float(0)
        _check(lib().nbgpu_last_kernel_ms(self.h, C.byref(ms)), self.h)
        return ms.value

    def launch_count(self):
        return int(lib().nbgpu_launch_count(self.h))

    def timer_begin(self):
        _check(lib().nbgpu_timer_begin(self.h), self.h)

    def timer_end(self):
        ms = C.c_float(0)
        _check(lib().nbgpu_timer_end(self.h, C.byref(ms)), self.h)
        return ms.value

    def geometry(self):
        g = np.zeros(8, np.int32)
        _check(lib().nbgpu_geometry(self.h, _i(g)), self.h)
        return dict(zip(("grid", "frames_per_cta", "cns_per_step", "steps_per_pass", "smem_bytes", "slots", "warps_per_cta",
                         "cns_per_warp"), [int(x) for x in g]))

    def slow_selects(self):
        return int(lib().nbgpu_slow_selects(self.h))

    def get_state(self, frame):
        app = np.zeros((self.code.N, self.code.q), np.float32)
        ctov = np.zeros((self.code.E, self.code.q), np.float32)
        _check(lib().nbgpu_get_state(self.h, frame, _f(app), _f(ctov)), self.h)
        return app, ctov

    def channel(self, noisy, sigma, want_sorted=False):
        noisy = np.ascontiguousarray(noisy, np.float32).reshape(-1, self.code.N, self.code.logq)
        B = noisy.shape[0]
        llr = np.zeros((B, self.code.N, self.code.q), np.float32)
        il = np.zeros_like(llr) if want_sorted else None
        ig = np.zeros(llr.shape, np.int32) if want_sorted else None
        _check(lib().nbgpu_channel_awgn_bpsk(self.h, _f(noisy), C.c_float(sigma), B, _f(llr), _f(il), _i(ig)), self.h)
        return (llr, il, ig) if want_sorted else llr

    def check_node(self, node, vllr, vgf):
        dc = int(self.code.row_deg[node])
        vllr = np.ascontiguousarray(vllr, np.float32).reshape(-1, dc, self.p.n_m)
        vgf = np.ascontiguousarray(vgf, np.int32).reshape(-1, dc, self.p.n_m)
        B = vllr.shape[0]
        cl = np.zeros((B, dc, self.code.q), np.float32)
        cg = np.zeros((B, dc, self.code.q), np.int32)
        _check(lib().nbgpu_check_node(self.h, node, _f(vllr), _i(vgf), _f(cl), _i(cg), B), self.h)
        return cl, cg

    def elementary_step(self, in1, in2, idx1, idx2):
        n_m = self.p.n_m
        in1 = np.ascontiguousarray(in1, np.float32).reshape(-1, n_m)
        in2 = np.ascontiguousarray(in2, np.float32).reshape(-1, n_m)
        idx1 = np.ascontiguousarray(idx1, np.int32).reshape(-1, n_m)
        idx2 = np.ascontiguousarray(idx2, np.int32).reshape(-1, n_m)
        B = in1.shape[0]
        out = np.zeros((B, n_m), np.float32)
        io = np.zeros((B, n_m), np.int32)
        _check(lib().nbgpu_elementary_step(self.h, _f(in1), _f(in2), _i(idx1), _i(idx2), _f(out), _i(io), B), self.h)
        return out, io

    def select_nm(self, rows):
        rows = np.ascontiguousarray(rows, np.float32).reshape(-1, self.code.q)
        B = rows.shape[0]
        ol = np.zeros((B, self.p.n_m), np.float32)
        og = np.zeros((B, self.p.n_m), np.int32)
        _check(lib().nbgpu_select_nm(self.h, _f(rows), _f(ol), _i(og), B), self.h)
        return ol, og

    def decision_syndrome(self, app):
        app = np.ascontiguousarray(app, np.float32).reshape(-1, self.code.N, self.code.q)
        B = app.shape[0]
        d = np.zeros((B, self.code.N), np.int32)
        s = np.zeros(B, np.int32)
        _check(lib().nbgpu_decision_syndrome(self.h, _f(app), _i(d), _i(s), B), self.h)
        return d, s

    def close(self):
        if self.h:
            lib().nbgpu_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def config_table(dc, d1, d2, d3, trunc=0):
    """nbgpu_config_table: the [size, dc] configuration table of the syndrome-based check node."""
    size = lib().nbgpu_config_table(dc, d1, d2, d3, trunc, None, 0)
    if size < 0:
        _check(size)
    t = np.zeros((size, dc), np.int32)
    _check(min(lib().nbgpu_config_table(dc, d1, d2, d3, trunc, _i(t), size), 0))
    return t


def pin(arr):
    """Page-lock a numpy array in place (nbgpu_host_register); returns the array."""
    _check(lib().nbgpu_host_register(C.c_void_p(arr.ctypes.data), arr.nbytes))
    return arr


def unpin(arr):
    _check(lib().nbgpu_host_unregister(C.c_void_p(arr.ctypes.data)))


from . import multigpu  # noqa: E402,F401  (frame sharding over ranks; host logic only)


def device_count():
    return int(lib().nbgpu_device_count())


def version():
    return lib().nbgpu_version().decode()
