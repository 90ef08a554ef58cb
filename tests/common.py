"""Shared helpers of the test-suite: fixture loading, matrix lookup, seeded frame generation."""
import glob
import hashlib
import os

import numpy as np
import pytest

import oracle_lib as ol

ROOT = ol.ROOT
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
MATRIX_DIRS = [os.path.join(ol.REF_DIR), "/root/reference"]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def matrix_path(rel):
    """rel = 'matrices/...' as in the reference tree; the copies made by oracle/Makefile travel to the GPU box."""
    for d in MATRIX_DIRS:
        p = os.path.join(d, rel)
        if os.path.exists(p):
            return p
    pytest.skip("matrix %s not available (build oracle/_ref where /root/reference exists)" % rel)


def golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
        self.z = z
        self.name = name
        self.frames, self.nb_iter_max, self.n_m, self.nb_oper = [int(x) for x in z["args"]]
        self.ebn = float(z["ebn"]); self.offset = float(z["offset"])
        self.matrix = str(z["matrix"]); self.dialect = 2 if str(z["dialect"]) == "kn" else 1
        self.N, self.M, self.GF, self.logGF, self.E = [int(x) for x in z["header"]]
        self.nf = z["nbin"].shape[0]
        self.npasses = z["npasses"]
        self.decide = z["decide"].astype(np.int32)
        self.synd = z["synd"]
        self.app_sha = z["app_sha"]
        self.llr_sha = z["llr_sha"]
        # (d1, d2, d3, trunc, n_cv) when the fixture was produced with the reference's syndrome_ems as check node
        self.synd_params = tuple(int(x) for x in z["synd_params"]) if "synd_params" in z else None

    def final(self, f, nb_iter_max=None):
        """(decide, synd, iters) the reference reports for frame f when run with nb_iter_max (<= the fixture's)."""
        P = (nb_iter_max or self.nb_iter_max) - 1
        p = min(int(self.npasses[f]), P)
        synd = int(self.synd[f, p - 1])
        iters = p if synd == 0 else P + 1          # sum_it += iter+1 (NB_LDPC.c:474): iter == P when never converged
        return self.decide[f, p - 1], synd, iters, p


def oracle_frames(o, nframes, ebn, want_llr=True):
    """The reference's frame stream (NB_LDPC.c:250-262) from the oracle port: per frame codeword, bits, noisy, llr."""
    o.prepare_encoder()
    o.rng_default()
    sigma = o.sigma(ebn)
    out = []
    for _ in range(nframes):
        cw, nbin = o.random_codeword()
        noisy = o.channel_noise(nbin, ebn)
        llr = o.channel_llr(noisy, sigma) if want_llr else None
        out.append(dict(cw=cw, nbin=nbin, noisy=noisy, llr=llr))
    return out, sigma


def product_frames(code, nframes, ebn):
    """Same stream from the product's host layer (nbgpu_random_codeword / nbgpu_awgn_bpsk_noise)."""
    code.prepare_encoder()
    code.rng_default()
    out = []
    for _ in range(nframes):
        cw, nbin = code.random_codeword()
        noisy = code.noise(nbin, ebn)
        out.append(dict(cw=cw, nbin=nbin, noisy=noisy))
    return out, code.sigma(ebn)


def random_regular_code(rng, N, M, q, dc, dv=2):
    """A random check-regular Tanner graph without repeated columns in a row (arrays for nbgpu_code_from_arrays)."""
    assert N * dv == M * dc
    for _ in range(200):
        slots = np.repeat(np.arange(N), dv)
        rng.shuffle(slots)
        rows = slots.reshape(M, dc)
        if all(len(set(r)) == dc for r in rows):
            break
    else:
        raise RuntimeError("could not build a simple graph")
    col = rows.reshape(-1).astype(np.int32)
    val = rng.integers(1, q, size=M * dc).astype(np.int32)
    return dict(N=N, M=M, q=q, row_deg=np.full(M, dc, np.int32), col=col, val=val)


def write_alist_ubs(path, a):
    with open(path, "w") as f:
        N, M, q = a["N"], a["M"], a["q"]
        f.write("%d %d %d\n" % (N, M, q))
        cd = np.bincount(a["col"], minlength=N)
        f.write(" ".join(str(x) for x in cd) + "\n")
        f.write(" ".join(str(x) for x in a["row_deg"]) + "\n")
        e = 0
        for m in range(M):
            f.write(" ".join(str(x) for x in a["col"][e:e + a["row_deg"][m]]) + "\n"); e += a["row_deg"][m]
        e = 0
        for m in range(M):
            f.write(" ".join(str(x) for x in a["val"][e:e + a["row_deg"][m]]) + "\n"); e += a["row_deg"][m]


def write_alist_full(path, a, pad=False):
    """full alist: max-degree line, column lists then row lists of (index 1-based, exponent) pairs (matrices/KN/N64800_*);
    pad=True fills every list with '0 0' pairs up to the maximum degree, as MacKay's tools do for irregular codes"""
    N, M, q = a["N"], a["M"], a["q"]
    rd = np.asarray(a["row_deg"]); col = np.asarray(a["col"]); val = np.asarray(a["val"])
    rp = np.concatenate([[0], np.cumsum(rd)])
    cols = [[] for _ in range(N)]
    for m in range(M):
        for e in range(rp[m], rp[m + 1]):
            cols[col[e]].append((m + 1, val[e] - 1))
    cd = [len(c) for c in cols]
    dv, dc = max(cd), int(rd.max())
    with open(path, "w") as f:
        f.write("%d %d %d\n%d %d\n" % (N, M, q, dv, dc))
        f.write(" ".join(map(str, cd)) + "\n" + " ".join(map(str, rd)) + "\n")
        for c in cols:
            items = c + ([(0, 0)] * (dv - len(c)) if pad else [])
            f.write("   ".join("%d %d" % it for it in items) + "\n")
        for m in range(M):
            items = [(col[e] + 1, val[e] - 1) for e in range(rp[m], rp[m + 1])]
            items += [(0, 0)] * (dc - len(items)) if pad else []
            f.write("   ".join("%d %d" % it for it in items) + "\n")


def write_alist_kn(path, a):
    """KN dialect (init.c:211-227): degrees, then per row (column 1-based, exponent) pairs"""
    N, M, q = a["N"], a["M"], a["q"]
    rd = np.asarray(a["row_deg"]); col = np.asarray(a["col"]); val = np.asarray(a["val"])
    rp = np.concatenate([[0], np.cumsum(rd)])
    with open(path, "w") as f:
        f.write("%d %d %d\n\n" % (N, M, q))
        f.write(" ".join(map(str, np.bincount(col, minlength=N))) + "\n" + " ".join(map(str, rd)) + "\n\n")
        for m in range(M):
            f.write("   ".join("%d %d" % (col[e] + 1, val[e] - 1) for e in range(rp[m], rp[m + 1])) + "\n")
