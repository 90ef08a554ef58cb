"""ctypes bindings for the CHECKERS used by the tests (never by the product):

* ``Oracle``  -> oracle/liboracle.so  (our plain-C restatement, oracle/nbldpc_oracle.c)
* ``RefShim`` -> oracle/_ref/libref.so (the unmodified reference objects + oracle/ref_shim.c)
* ``read_trace`` -> parser for the traces written by oracle/_ref/essai_probe (oracle/ref_probes.c)
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
REFERENCE_SRC = "/root/reference"

c_int_p = C.POINTER(C.c_int)
c_float_p = C.POINTER(C.c_float)


def _ip(a):
    return a.ctypes.data_as(c_int_p)


def _fp(a):
    return a.ctypes.data_as(c_float_p)


def build_oracle():
    """Compile oracle/liboracle.so (and oracle/_ref when the reference sources are present)."""
    subprocess.run(["make", "-s", "port"], cwd=ORACLE_DIR, check=True)
    if os.path.isdir(REFERENCE_SRC):
        subprocess.run(["make", "-s", "ref"], cwd=ORACLE_DIR, check=True)


def have_ref():
    return os.path.exists(os.path.join(REF_DIR, "libref.so"))


class NboCode(C.Structure):
    _fields_ = [("N", C.c_int), ("M", C.c_int), ("K", C.c_int), ("GF", C.c_int), ("logGF", C.c_int),
                ("E", C.c_int), ("dc_max", C.c_int), ("rate", C.c_float),
                ("row_deg", c_int_p), ("row_ptr", c_int_p), ("col", c_int_p), ("val", c_int_p),
                ("bingf", c_int_p), ("addgf", c_int_p), ("mulgf", c_int_p), ("divgf", c_int_p),
                ("matUT", c_int_p), ("perm", c_int_p)]


class NboParams(C.Structure):
    _fields_ = [("n_m", C.c_int), ("nb_oper", C.c_int), ("nb_iter_max", C.c_int), ("offset", C.c_float),
                ("ecn", C.c_int), ("force_passes", C.c_int), ("n_cv", C.c_int), ("cfg_size", C.c_int),
                ("cfg", c_int_p)]


class NboRng(C.Structure):
    _fields_ = [("x", C.c_uint64)]


class Oracle:
    """Thin wrapper over liboracle.so for one code."""
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            path = os.path.join(ORACLE_DIR, "liboracle.so")
            if not os.path.exists(path):
                build_oracle()
            L = C.CDLL(path)
            L.nbo_load.restype = C.POINTER(NboCode)
            L.nbo_load.argtypes = [C.c_char_p, C.c_int]
            L.nbo_free.argtypes = [C.POINTER(NboCode)]
            L.nbo_prepare_encoder.argtypes = [C.POINTER(NboCode)]
            L.nbo_drand48.restype = C.c_double
            L.nbo_sigma.restype = C.c_float
            L.nbo_sigma.argtypes = [C.POINTER(NboCode), C.c_float]
            L.nbo_rng_skip.argtypes = [C.POINTER(NboRng), C.c_uint64]
            L.nbo_channel_noise.argtypes = [C.POINTER(NboCode), C.POINTER(NboRng), c_int_p, C.c_float, c_float_p]
            L.nbo_channel_llr.argtypes = [C.POINTER(NboCode), c_float_p, C.c_float, c_float_p]
            L.nbo_apsk64_table.argtypes = [c_float_p]
            L.nbo_sigma_apsk64.restype = C.c_float
            L.nbo_sigma_apsk64.argtypes = [C.c_float]
            L.nbo_channel_noise_apsk64.argtypes = [C.POINTER(NboCode), C.POINTER(NboRng), c_int_p, C.c_float, c_float_p]
            L.nbo_channel_llr_apsk64.argtypes = [C.POINTER(NboCode), c_float_p, C.c_float, c_float_p]
            L.nbo_check_node_bubble.argtypes = [C.POINTER(NboCode), C.c_int, c_float_p, c_int_p, c_float_p, c_int_p,
                                                C.c_int, C.c_int, C.c_float]
            L.nbo_check_node_syndrome.argtypes = [C.POINTER(NboCode), C.c_int, c_float_p, c_int_p, c_float_p, c_int_p,
                                                  C.c_int, c_int_p, C.c_int, C.c_float, C.c_int]
            L.nbo_build_config_table.restype = c_int_p
            L.nbo_build_config_table.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_int_p]
            L.nbo_monte_carlo.argtypes = [C.POINTER(NboCode), C.POINTER(NboParams), C.c_int, C.c_float,
                                          C.POINTER(C.c_long)]
            cls._lib = L
        return cls._lib

    def __init__(self, path, dialect=1):
        L = self.lib()
        self.h = L.nbo_load(path.encode(), dialect)
        if not self.h:
            raise IOError("oracle: cannot load %s" % path)
        c = self.h.contents
        self.N, self.M, self.K, self.GF, self.logGF, self.E, self.dc_max = (
            c.N, c.M, c.K, c.GF, c.logGF, c.E, c.dc_max)
        self.rate = c.rate
        self.row_deg = np.ctypeslib.as_array(c.row_deg, (self.M,)).copy()
        self.row_ptr = np.ctypeslib.as_array(c.row_ptr, (self.M + 1,)).copy()
        self.col = np.ctypeslib.as_array(c.col, (self.E,)).copy()
        self.val = np.ctypeslib.as_array(c.val, (self.E,)).copy()
        self.bingf = np.ctypeslib.as_array(c.bingf, (self.GF, self.logGF)).copy()
        self.addgf = np.ctypeslib.as_array(c.addgf, (self.GF, self.GF)).copy()
        self.mulgf = np.ctypeslib.as_array(c.mulgf, (self.GF, self.GF)).copy()
        self.divgf = np.ctypeslib.as_array(c.divgf, (self.GF, self.GF)).copy()
        self.rng = NboRng()
        L.nbo_rng_default(C.byref(self.rng))

    def close(self):
        if self.h:
            self.lib().nbo_free(self.h)
            self.h = None

    # ---- frame source / channel -------------------------------------------------------------
    def rng_default(self):
        self.lib().nbo_rng_default(C.byref(self.rng))

    def rng_skip(self, n):
        self.lib().nbo_rng_skip(C.byref(self.rng), n)

    def drand48(self):
        return self.lib().nbo_drand48(C.byref(self.rng))

    def prepare_encoder(self):
        if self.lib().nbo_prepare_encoder(self.h) != 0:
            raise ValueError("H is rank deficient")

    def random_codeword(self):
        cw = np.zeros(self.N, np.int32)
        nbin = np.zeros((self.N, self.logGF), np.int32)
        self.lib().nbo_random_codeword(self.h, C.byref(self.rng), _ip(cw), _ip(nbin))
        return cw, nbin

    def sigma(self, ebn):
        return self.lib().nbo_sigma(self.h, C.c_float(ebn))

    def channel_noise(self, nbin, ebn):
        nbin = np.ascontiguousarray(nbin, np.int32)
        noisy = np.zeros((self.N, self.logGF), np.float32)
        self.lib().nbo_channel_noise(self.h, C.byref(self.rng), _ip(nbin), C.c_float(ebn), _fp(noisy))
        return noisy

    def channel_llr(self, noisy, sigma):
        noisy = np.ascontiguousarray(noisy, np.float32)
        llr = np.zeros((self.N, self.GF), np.float32)
        self.lib().nbo_channel_llr(self.h, _fp(noisy), C.c_float(sigma), _fp(llr))
        return llr

    # ---- 64-APSK channel (ModelChannel_AWGN_64), GF(64) codes ----
    def apsk64_table(self):
        mod = np.zeros((64, 2), np.float32)
        self.lib().nbo_apsk64_table(_fp(mod))
        return mod

    def sigma_apsk64(self, ebn):
        return self.lib().nbo_sigma_apsk64(C.c_float(ebn))

    def channel_noise_apsk64(self, nbin, ebn):
        nbin = np.ascontiguousarray(nbin, np.int32)
        noisy = np.zeros((self.N, 2), np.float32)
        self.lib().nbo_channel_noise_apsk64(self.h, C.byref(self.rng), _ip(nbin), C.c_float(ebn), _fp(noisy))
        return noisy

    def channel_llr_apsk64(self, noisy, sigma):
        noisy = np.ascontiguousarray(noisy, np.float32)
        llr = np.zeros((self.N, self.GF), np.float32)
        self.lib().nbo_channel_llr_apsk64(self.h, _fp(noisy), C.c_float(sigma), _fp(llr))
        return llr

    def sort_intrinsic(self, llr):
        llr = np.ascontiguousarray(llr, np.float32)
        il = np.zeros((self.N, self.GF), np.float32)
        ig = np.zeros((self.N, self.GF), np.int32)
        self.lib().nbo_sort_intrinsic(self.h, _fp(llr), _fp(il), _ip(ig))
        return il, ig

    # ---- decoder pieces ------------------------------------------------------------------------
    def select_nm(self, row, n_m):
        row = np.ascontiguousarray(row, np.float32)
        ol = np.zeros(n_m, np.float32)
        og = np.zeros(n_m, np.int32)
        self.lib().nbo_select_nm(_fp(row), C.c_int(len(row)), C.c_int(n_m), _fp(ol), _ip(og))
        return ol, og

    def elementary_step(self, in1, in2, idx1, idx2, n_m, nb_oper):
        in1 = np.ascontiguousarray(in1, np.float32); in2 = np.ascontiguousarray(in2, np.float32)
        idx1 = np.ascontiguousarray(idx1, np.int32); idx2 = np.ascontiguousarray(idx2, np.int32)
        out = np.zeros(n_m, np.float32); io = np.zeros(n_m, np.int32)
        add = np.ascontiguousarray(self.addgf, np.int32)
        self.lib().nbo_elementary_step(_fp(in1), _fp(in2), _ip(idx1), _ip(idx2), _fp(out), _ip(io), _ip(add),
                                       C.c_int(self.GF), C.c_int(n_m), C.c_int(nb_oper))
        return out, io

    def check_node(self, node, vllr, vgf, n_m, nb_oper, offset):
        dc = int(self.row_deg[node])
        vllr = np.ascontiguousarray(vllr, np.float32).reshape(dc, n_m)
        vgf = np.ascontiguousarray(vgf, np.int32).reshape(dc, n_m)
        cl = np.zeros((dc, self.GF), np.float32); cg = np.zeros((dc, self.GF), np.int32)
        self.lib().nbo_check_node_bubble(self.h, node, _fp(vllr), _ip(vgf), _fp(cl), _ip(cg), n_m, nb_oper,
                                         C.c_float(offset))
        return cl, cg

    def build_config_table(self, dc, d1, d2, d3, trunc):
        size = C.c_int(0)
        p = self.lib().nbo_build_config_table(dc, d1, d2, d3, trunc, C.byref(size))
        return np.ctypeslib.as_array(p, (size.value, dc)).copy()

    def check_node_syndrome(self, node, vllr, vgf, n_m, cfg, offset, n_cv):
        dc = int(self.row_deg[node])
        vllr = np.ascontiguousarray(vllr, np.float32).reshape(dc, n_m)
        vgf = np.ascontiguousarray(vgf, np.int32).reshape(dc, n_m)
        cfg = np.ascontiguousarray(cfg, np.int32)
        cl = np.zeros((dc, self.GF), np.float32); cg = np.zeros((dc, self.GF), np.int32)
        self.lib().nbo_check_node_syndrome(self.h, node, _fp(vllr), _ip(vgf), _fp(cl), _ip(cg), n_m, _ip(cfg),
                                           cfg.shape[0], C.c_float(offset), n_cv)
        return cl, cg

    def decision(self, app):
        app = np.ascontiguousarray(app, np.float32)
        d = np.zeros(self.N, np.int32)
        self.lib().nbo_decision(_fp(app), self.N, self.GF, _ip(d))
        return d

    def syndrome(self, decide):
        decide = np.ascontiguousarray(decide, np.int32)
        return self.lib().nbo_syndrome(self.h, _ip(decide))

    def params(self, n_m, nb_oper, nb_iter_max, offset, force=False, ecn=0, cfg=None, n_cv=0):
        p = NboParams()
        p.n_m, p.nb_oper, p.nb_iter_max, p.offset = n_m, nb_oper, nb_iter_max, offset
        p.ecn, p.force_passes, p.n_cv = ecn, int(force), n_cv
        if cfg is not None:
            self._cfg_keep = np.ascontiguousarray(cfg, np.int32)
            p.cfg = _ip(self._cfg_keep)
            p.cfg_size = self._cfg_keep.shape[0]
        return p

    def decode_frame(self, llr, n_m, nb_oper, nb_iter_max, offset, force=False, want_state=False,
                     ecn=0, cfg=None, n_cv=0):
        """Returns dict(decide, synd, iters, passes, decide_trace, synd_trace[, app, ctov])."""
        llr = np.ascontiguousarray(llr, np.float32)
        p = self.params(n_m, nb_oper, nb_iter_max, offset, force, ecn, cfg, n_cv)
        P = max(nb_iter_max - 1, 0)
        decide = np.zeros(self.N, np.int32)
        dtr = np.zeros((max(P, 1), self.N), np.int32)
        strc = np.full(max(P, 1), -1, np.int32)
        synd = C.c_int(0); iters = C.c_int(0)
        app = np.zeros((self.N, self.GF), np.float32) if want_state else None
        ctov = np.zeros((self.E, self.GF), np.float32) if want_state else None
        passes = self.lib().nbo_decode_frame(self.h, C.byref(p), _fp(llr), _ip(decide), C.byref(synd),
                                             C.byref(iters), _ip(dtr), _ip(strc),
                                             _fp(app) if want_state else None, _fp(ctov) if want_state else None)
        r = dict(decide=decide, synd=synd.value, iters=iters.value, passes=passes,
                 decide_trace=dtr[:passes], synd_trace=strc[:passes])
        if want_state:
            r["app"] = app; r["ctov"] = ctov
        return r

    def monte_carlo(self, frames, ebn, n_m, nb_oper, nb_iter_max, offset):
        p = self.params(n_m, nb_oper, nb_iter_max, offset)
        stats = (C.c_long * 6)()
        self.lib().nbo_monte_carlo(self.h, C.byref(p), frames, C.c_float(ebn), stats)
        return list(stats)


class RefShim:
    """The reference's own functions (oracle/_ref/libref.so)."""
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(os.path.join(REF_DIR, "libref.so"))
            L.refshim_open.restype = C.c_void_p
            L.refshim_open.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int]
            L.refshim_rate.restype = C.c_float
            L.refshim_rate.argtypes = [C.c_void_p]
            for n in ("refshim_info", "refshim_graph", "refshim_tables", "refshim_random_codeword",
                      "refshim_channel_bpsk", "refshim_channel_apsk64", "refshim_elementary_step", "refshim_check_node",
                      "refshim_decision_syndrome", "refshim_build_config", "refshim_get_config",
                      "refshim_syndrome_ems"):
                getattr(L, n).argtypes = None
            cls._lib = L
        return cls._lib

    def __init__(self, path, kn=False, n_m=20, encoder=False):
        L = self.lib()
        self.h = C.c_void_p(L.refshim_open(path.encode(), int(kn), n_m, int(encoder)))
        if not self.h:
            raise IOError("reference: cannot load %s" % path)
        info = np.zeros(8, np.int32)
        L.refshim_info(self.h, _ip(info))
        self.N, self.M, self.GF, self.logGF, self.E, self.n_m, self.dc0, self.K = [int(x) for x in info]
        self.row_deg = np.zeros(self.M, np.int32)
        self.col = np.zeros(self.E, np.int32); self.val = np.zeros(self.E, np.int32)
        L.refshim_graph(self.h, _ip(self.row_deg), _ip(self.col), _ip(self.val))
        self.bingf = np.zeros((self.GF, self.logGF), np.int32)
        self.addgf = np.zeros((self.GF, self.GF), np.int32)
        self.mulgf = np.zeros((self.GF, self.GF), np.int32)
        self.divgf = np.zeros((self.GF, self.GF), np.int32)
        L.refshim_tables(self.h, _ip(self.bingf), _ip(self.addgf), _ip(self.mulgf), _ip(self.divgf))
        self.rate = L.refshim_rate(self.h)

    def seed_default(self):
        self.lib().refshim_seed_default()

    def random_codeword(self):
        cw = np.zeros(self.N, np.int32); nbin = np.zeros((self.N, self.logGF), np.int32)
        assert self.lib().refshim_random_codeword(self.h, _ip(cw), _ip(nbin)) == 0
        return cw, nbin

    def channel(self, nbin, ebn):
        nbin = np.ascontiguousarray(nbin, np.int32)
        il = np.zeros((self.N, self.GF), np.float32); ig = np.zeros((self.N, self.GF), np.int32)
        self.lib().refshim_channel_bpsk(self.h, _ip(nbin), C.c_float(ebn), _fp(il), _ip(ig))
        return il, ig

    def channel_apsk64(self, nbin, ebn):
        """ModelChannel_AWGN_64 of the reference on the process-wide drand48 state: sorted intrinsic LLR / GF"""
        nbin = np.ascontiguousarray(nbin, np.int32)
        il = np.zeros((self.N, self.GF), np.float32); ig = np.zeros((self.N, self.GF), np.int32)
        self.lib().refshim_channel_apsk64(self.h, _ip(nbin), C.c_float(ebn), _fp(il), _ip(ig))
        return il, ig

    def elementary_step(self, in1, in2, idx1, idx2, n_m, nb_oper):
        in1 = np.ascontiguousarray(in1, np.float32); in2 = np.ascontiguousarray(in2, np.float32)
        idx1 = np.ascontiguousarray(idx1, np.int32); idx2 = np.ascontiguousarray(idx2, np.int32)
        out = np.zeros(n_m, np.float32); io = np.zeros(n_m, np.int32)
        self.lib().refshim_elementary_step(self.h, _fp(in1), _fp(in2), _ip(idx1), _ip(idx2), _fp(out), _ip(io),
                                           C.c_int(n_m), C.c_int(nb_oper))
        return out, io

    def check_node(self, node, vllr, vgf, nb_oper, offset):
        dc = int(self.row_deg[node])
        vllr = np.ascontiguousarray(vllr, np.float32); vgf = np.ascontiguousarray(vgf, np.int32)
        cl = np.zeros((dc, self.GF), np.float32); cg = np.zeros((dc, self.GF), np.int32)
        self.lib().refshim_check_node(self.h, C.c_int(node), _fp(vllr), _ip(vgf), _fp(cl), _ip(cg),
                                      C.c_int(nb_oper), C.c_float(offset))
        return cl, cg

    def decision_syndrome(self, app):
        app = np.ascontiguousarray(app, np.float32)
        d = np.zeros(self.N, np.int32)
        s = self.lib().refshim_decision_syndrome(self.h, _fp(app), _ip(d))
        return d, s

    def build_config(self, dc, d1, d2, d3, trunc):
        size = self.lib().refshim_build_config(self.h, C.c_int(dc), C.c_int(d1), C.c_int(d2), C.c_int(d3),
                                               C.c_int(trunc))
        out = np.zeros((size, dc), np.int32)
        self.lib().refshim_get_config(self.h, C.c_int(dc), _ip(out))
        return out

    def syndrome_ems(self, node, vllr, vgf, dc, offset, n_cv):
        vllr = np.ascontiguousarray(vllr, np.float32); vgf = np.ascontiguousarray(vgf, np.int32)
        cl = np.zeros((dc, self.GF), np.float32); cg = np.zeros((dc, self.GF), np.int32)
        self.lib().refshim_syndrome_ems(self.h, C.c_int(node), _fp(vllr), _ip(vgf), _fp(cl), _ip(cg),
                                        C.c_int(dc), C.c_float(offset), C.c_int(n_cv))
        return cl, cg


def run_probe(args, trace=None, level=1, force=False, dialect="ubs", summary=None, cwd=None, synd=None):
    """Run the interposed reference binary; args = [frames, iters, matrix, EbN, n_m, offset, nbOper]."""
    env = dict(os.environ)
    if trace:
        env["NBREF_TRACE"] = trace
    env["NBREF_LEVEL"] = str(level)
    env["NBREF_FORCE"] = "1" if force else "0"
    env["NBREF_DIALECT"] = dialect
    if summary:
        env["NBREF_SUMMARY"] = summary
    env.pop("NBREF_ECN", None)
    if synd:                                   # (d1, d2, d3, trunc, n_cv): run the reference's syndrome_ems as check node
        env["NBREF_ECN"] = "syndrome"
        env["NBREF_SYND"] = ",".join(str(int(x)) for x in synd)
    cwd = cwd or REF_DIR
    os.makedirs(os.path.join(cwd, "data"), exist_ok=True)
    r = subprocess.run([os.path.join(REF_DIR, "essai_probe")] + [str(a) for a in args], cwd=cwd, env=env,
                       stdin=subprocess.DEVNULL, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, check=True)
    return r.stdout.decode(errors="replace")


def read_trace(path):
    """Parse an essai_probe trace into dict(header=..., frames=[{nbin, illr, igf, passes:[{synd, decide, app}], cn:[...]}])."""
    data = open(path, "rb").read()
    off = 0
    hdr = None
    frames = []
    while off < len(data):
        tag, nb = np.frombuffer(data, np.int32, 2, off)
        off += 8
        pay = data[off:off + nb]
        off += nb
        if tag == 1:
            h = np.frombuffer(pay, np.int32)
            hdr = dict(N=int(h[0]), M=int(h[1]), GF=int(h[2]), logGF=int(h[3]), n_m=int(h[4]), E=int(h[5]),
                       dc0=int(h[6]))
        elif tag == 2:
            frames.append(dict(nbin=np.frombuffer(pay, np.int8).reshape(hdr["N"], hdr["logGF"]).astype(np.int32),
                               passes=[], cn=[]))
        elif tag == 3:
            cnt = hdr["N"] * hdr["GF"]
            frames[-1]["illr"] = np.frombuffer(pay, np.float32, cnt).reshape(hdr["N"], hdr["GF"])
            frames[-1]["igf"] = np.frombuffer(pay, np.int16, cnt, cnt * 4).reshape(hdr["N"], hdr["GF"]).astype(np.int32)
        elif tag == 4:
            frames[-1]["passes"].append(dict(synd=int(np.frombuffer(pay, np.int32, 1)[0]),
                                             decide=np.frombuffer(pay, np.int16, hdr["N"], 4).astype(np.int32)))
        elif tag == 5:
            frames[-1]["passes"][-1]["app"] = np.frombuffer(pay, np.float32).reshape(hdr["N"], hdr["GF"])
        elif tag == 6:
            node, dc = np.frombuffer(pay, np.int32, 2)
            nm, GF = hdr["n_m"], hdr["GF"]
            o = 8
            il = np.frombuffer(pay, np.float32, dc * nm, o); o += dc * nm * 4
            ig = np.frombuffer(pay, np.int16, dc * nm, o); o += dc * nm * 2
            ol = np.frombuffer(pay, np.float32, dc * GF, o); o += dc * GF * 4
            og = np.frombuffer(pay, np.int16, dc * GF, o)
            frames[-1]["cn"].append(dict(node=int(node), in_llr=il.reshape(dc, nm), in_gf=ig.reshape(dc, nm).astype(np.int32),
                                         out_llr=ol.reshape(dc, GF), out_gf=og.reshape(dc, GF).astype(np.int32)))
    return dict(header=hdr, frames=frames)


def dense_from_intrinsic(illr, igf):
    """NB_LDPC.c:281-288: APP[n][intrinsic_GF[n][k]] = intrinsic_LLR[n][k]."""
    N, GF = illr.shape
    out = np.zeros((N, GF), np.float32)
    np.put_along_axis(out, igf.astype(np.int64), illr, axis=1)
    return out
