#!/usr/bin/env python3
"""Generate the golden fixtures in tests/golden/ from the UNMODIFIED reference.

Needs oracle/_ref (built by `make -C oracle ref` from /root/reference; see oracle/Makefile).  Each
fixture is one run of the reference's own main() (oracle/_ref/essai_probe = NB_LDPC.c with
Decision/Syndrom/ModelChannel_AWGN_BPSK/CheckPassLogEMS interposed by oracle/ref_probes.c), i.e. the
reference's decode loop NB_LDPC.c:250-511 on its own drand48 stream.  Stored per frame:

    nbin            codeword bits handed to the channel (int8 [N, logGF])
    llr_sha / llr   sha256 (and, for small codes, the values) of the dense channel LLR in GF order
    decide, synd    hard decisions and Syndrom value after every executed pass
    app_sha         sha256 of APP[N][GF] after every executed pass
    cn_*            (small codes, first frames) inputs/outputs of every CheckPassLogEMS call
    stdout_tail     the console statistics line

Usage: python tests/golden/make_golden.py            (rewrites tests/golden/*.npz)
"""
import hashlib
import os
import re
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle_lib as ol  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# name, frames, NbIterMax, matrix, EbN, n_m, offset, NbOper, trace level, keep_llr, keep_cn_frames, dialect
CASES = [
    ("n96_gf64_nm20", 24, 10, "matrices/N96_K48_GF64", 2.0, 20, 0.3, 25, 3, True, 2, "ubs"),
    ("n96_gf64_kn_nm20", 6, 10, "matrices/KN/N96_K48_GF64.txt", 2.0, 20, 0.3, 25, 2, False, 0, "kn"),
    ("n96_gf256_kn_nm20", 4, 10, "matrices/KN/N96_K48_GF256.txt", 2.5, 20, 0.3, 25, 3, True, 1, "kn"),
    ("mat24_n48_gf64_nm8", 8, 6, "matrices/Mat24_N48_M24", 2.0, 8, 0.5, 12, 3, True, 1, "ubs"),
    ("mat28_n72_nm12", 6, 10, "matrices/Mat28_N72_M18", 2.4, 12, 0.3, 25, 3, True, 1, "ubs"),
    ("mat212_n96_nm16", 4, 10, "matrices/Mat212_N96_M16", 3.0, 16, 0.3, 25, 2, False, 0, "ubs"),
    ("mat24_n480_nm16", 6, 10, "matrices/Mat24_N480_M240", 1.5, 16, 0.3, 25, 2, False, 0, "ubs"),
    ("declercq_r12_gf64_nm20", 2, 10, "matrices/MatDeclercq_R12_GF64", 1.2, 20, 0.3, 25, 2, False, 0, "ubs"),
    ("ahmed_r34_gf16_nm16", 2, 10, "matrices/Ahmed_64800_R34_GF16", 3.0, 16, 0.3, 25, 2, False, 0, "ubs"),
    ("ad_r12_gf256_nm20", 2, 10, "matrices/AD_64800_R12_GF256", 2.0, 20, 0.3, 25, 2, False, 0, "ubs"),
    # check node = the reference's syndrome_ems (NB_LDPC.c:388 instead of :392); last field = (d1, d2, d3, trunc, n_cv)
    ("n96_gf64_nm20_synd", 12, 10, "matrices/N96_K48_GF64", 2.0, 20, 0.3, 25, 3, True, 1, "ubs", (19, 15, 5, 1000, 25)),
    ("mat24_n480_nm20_synd", 4, 10, "matrices/Mat24_N480_M240", 1.5, 20, 0.3, 25, 2, False, 0, "ubs", (19, 15, 5, 1000, 25)),
    ("mat24_n480_nm16_synd_small", 3, 10, "matrices/Mat24_N480_M240", 1.5, 16, 0.3, 25, 2, False, 0, "ubs", (15, 6, 3, 150, 12)),
    ("ad_r12_gf256_nm20_synd", 3, 10, "matrices/AD_64800_R12_GF256", 2.0, 20, 0.3, 25, 2, False, 0, "ubs", (19, 15, 5, 1000, 25)),
]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def make(case):
    name, frames, iters, matrix, ebn, n_m, offset, nbop, level, keep_llr, keep_cn, dialect = case[:12]
    synd = case[12] if len(case) > 12 else None
    with tempfile.TemporaryDirectory() as td:
        tr = os.path.join(td, "trace.bin")
        out = ol.run_probe([frames, iters, matrix, ebn, n_m, offset, nbop], trace=tr, level=level, dialect=dialect, synd=synd)
        t = ol.read_trace(tr)
    h = t["header"]
    fr = t["frames"]
    P = iters - 1
    N = h["N"]
    d = dict(args=np.array([frames, iters, n_m, nbop], np.int32), ebn=np.float32(ebn), offset=np.float32(offset),
             matrix=matrix, dialect=dialect, header=np.array([h["N"], h["M"], h["GF"], h["logGF"], h["E"]], np.int32))
    if synd:
        d["synd_params"] = np.array(synd, np.int32)
    nf = len(fr)
    d["nbin"] = np.stack([f["nbin"] for f in fr]).astype(np.int8)
    d["npasses"] = np.array([len(f["passes"]) for f in fr], np.int32)
    dec = np.full((nf, P, N), -1, np.int16)
    syn = np.full((nf, P), -1, np.int32)
    app_sha = np.full((nf, P), "", dtype="U64")
    llr_sha = []
    llrs = []
    for i, f in enumerate(fr):
        dense = ol.dense_from_intrinsic(f["illr"], f["igf"])
        llr_sha.append(sha(dense))
        if keep_llr:
            llrs.append(dense)
        for p, ps in enumerate(f["passes"]):
            dec[i, p] = ps["decide"]
            syn[i, p] = ps["synd"]
            if "app" in ps:
                app_sha[i, p] = sha(ps["app"])
    d["decide"] = dec
    d["synd"] = syn
    d["app_sha"] = app_sha
    d["llr_sha"] = np.array(llr_sha)
    d["illr_sha"] = np.array([sha(f["illr"]) for f in fr])
    d["igf_sha"] = np.array([sha(f["igf"].astype(np.int32)) for f in fr])
    if keep_llr:
        d["llr"] = np.stack(llrs)
    if keep_cn:
        cn_in_l, cn_in_g, cn_out_l, cn_node, cn_out_g = [], [], [], [], []
        for f in fr[:keep_cn]:
            for c in f["cn"]:
                cn_node.append(c["node"]); cn_in_l.append(c["in_llr"]); cn_in_g.append(c["in_gf"])
                cn_out_l.append(c["out_llr"])
                if synd:
                    cn_out_g.append(c["out_gf"])
                else:
                    assert (c["out_gf"] == np.arange(h["GF"])[None, :]).all()
        d["cn_node"] = np.array(cn_node, np.int32)
        d["cn_in_llr"] = np.stack(cn_in_l)
        d["cn_in_gf"] = np.stack(cn_in_g).astype(np.int16)
        d["cn_out_llr"] = np.stack(cn_out_l)
        if synd:
            d["cn_out_gf"] = np.stack(cn_out_g).astype(np.int16)
    m = re.findall(r"<\d+> FER=\s*(\d+)\s*/\s*(\d+)\s*=\s*[\d.]+\s*BER=\s*(\d+)\s*/\s*x\s*=\s*[\d.eE+-]+\s*avr_it=([\d.]+)", out)
    d["console"] = np.array([float(x) for x in m[-1]]) if m else np.zeros(4)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **d)
    print("%-28s frames=%d passes=%s console=%s  %.1f KB" % (name, nf, d["npasses"].tolist(), d["console"].tolist(),
                                                           os.path.getsize(path) / 1024))


def make_apsk64():
    """tests/golden/channels/apsk64_n96_gf64.npz: the reference's ModelChannel_AWGN_64 (channel.c:112-312) through
    oracle/_ref/libref.so on the unseeded drand48 stream: per frame the codeword bits and the sorted intrinsic LLR / GF."""
    from common import matrix_path
    r = ol.RefShim(matrix_path("matrices/N96_K48_GF64"), n_m=20, encoder=True)
    r.seed_default()
    ebn = 9.0
    nbin, illr, igf = [], [], []
    for _ in range(6):
        _, nb = r.random_codeword()
        il, ig = r.channel_apsk64(nb, ebn)
        nbin.append(nb); illr.append(il); igf.append(ig)
    os.makedirs(os.path.join(HERE, "channels"), exist_ok=True)
    path = os.path.join(HERE, "channels", "apsk64_n96_gf64.npz")
    np.savez_compressed(path, matrix="matrices/N96_K48_GF64", ebn=np.float32(ebn), nbin=np.stack(nbin).astype(np.int8),
                        illr=np.stack(illr), igf=np.stack(igf).astype(np.int16))
    print("channels/apsk64_n96_gf64   frames=6  %.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    if not ol.have_ref():
        ol.build_oracle()
    sel = sys.argv[1:]
    for c in CASES:
        if not sel or c[0] in sel:
            make(c)
    if not sel or "apsk64" in sel:
        make_apsk64()
