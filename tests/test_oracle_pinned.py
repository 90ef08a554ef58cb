"""The oracle (oracle/nbldpc_oracle.c) is only trustworthy once it is pinned:
 * against the golden fixtures generated from the unmodified reference (always), and
 * against the reference objects themselves (oracle/_ref/libref.so) when they are present.
The reference ships no tests or golden vectors of its own."""
import os

import numpy as np
import pytest

import oracle_lib as ol
from common import Golden, golden_names, matrix_path, oracle_frames, sha

needs_ref = pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built")

SMALL = [n for n in golden_names() if not n.startswith(("declercq", "ahmed", "ad_"))]
LARGE = [n for n in golden_names() if n.startswith(("declercq", "ahmed", "ad_"))]


@pytest.mark.parametrize("name", SMALL + LARGE)
def test_oracle_reproduces_reference_run(name):
    """Frame source, channel LLRs and the whole decode loop: per-pass decisions, syndromes and APP bits."""
    g = Golden(name)
    o = ol.Oracle(matrix_path(g.matrix), g.dialect)
    assert (o.N, o.M, o.GF, o.E) == (g.N, g.M, g.GF, g.E)
    nf = g.nf if name in SMALL else 1
    frames, sigma = oracle_frames(o, nf, g.ebn)
    for f, fr in enumerate(frames):
        assert (fr["nbin"] == g.z["nbin"][f]).all(), "codeword bits differ (RNG / encoder)"
        assert sha(fr["llr"]) == g.llr_sha[f], "channel LLR differs"
        il, ig = o.sort_intrinsic(fr["llr"])
        assert sha(il) == g.z["illr_sha"][f] and sha(ig) == g.z["igf_sha"][f]
        kw = {}
        if g.synd_params:                     # fixture produced with the reference's syndrome_ems as check node
            d1, d2, d3, trunc, n_cv = g.synd_params
            kw = dict(ecn=1, cfg=o.build_config_table(int(o.row_deg[0]), d1, d2, d3, trunc), n_cv=n_cv)
        r = o.decode_frame(fr["llr"], g.n_m, g.nb_oper, g.nb_iter_max, g.offset, want_state=True, **kw)
        np_ = int(g.npasses[f])
        assert r["passes"] == np_
        assert (r["decide_trace"] == g.decide[f, :np_]).all()
        assert (r["synd_trace"] == g.synd[f, :np_]).all()
        dec, synd, iters, _ = g.final(f)
        assert (r["decide"] == dec).all() and r["synd"] == synd and r["iters"] == iters
        if g.app_sha[f, np_ - 1]:
            assert sha(r["app"]) == g.app_sha[f, np_ - 1], "APP after the last pass is not bit-identical"
    o.close()


@pytest.mark.parametrize("name", [n for n in golden_names() if "cn_node" in Golden(n).z])
def test_oracle_check_node_on_recorded_messages(name):
    g = Golden(name)
    o = ol.Oracle(matrix_path(g.matrix), g.dialect)
    z = g.z
    cfg = None
    if g.synd_params:
        d1, d2, d3, trunc, n_cv = g.synd_params
        cfg = o.build_config_table(int(o.row_deg[0]), d1, d2, d3, trunc)
    for i in range(len(z["cn_node"])):
        if cfg is not None:
            cl, cg = o.check_node_syndrome(int(z["cn_node"][i]), z["cn_in_llr"][i], z["cn_in_gf"][i].astype(np.int32), g.n_m, cfg,
                                           g.offset, n_cv)
            assert cl.tobytes() == z["cn_out_llr"][i].tobytes()
            assert (cg == z["cn_out_gf"][i]).all()
            continue
        cl, cg = o.check_node(int(z["cn_node"][i]), z["cn_in_llr"][i], z["cn_in_gf"][i].astype(np.int32), g.n_m, g.nb_oper,
                              g.offset)
        assert cl.tobytes() == z["cn_out_llr"][i].tobytes()
        assert (cg == np.arange(g.GF)[None, :]).all()
    o.close()


def test_oracle_monte_carlo_known_answers():
    """Console lines recorded from the stock binary (SURVEY.md section 8c / BASELINE.md section 2)."""
    o = ol.Oracle(matrix_path("matrices/N96_K48_GF64"))
    s = o.monte_carlo(2000, 3.0, 20, 25, 10, 0.3)
    assert s[:4] == [2000, 24, s[2], 153] and s[5] == 2001
    assert "%.2f" % (s[4] / 2000.0) == "1.56"
    o.close()
    o = ol.Oracle(matrix_path("matrices/Mat24_N480_M240"))
    s = o.monte_carlo(60, 1.5, 16, 25, 10, 0.3)
    g = Golden("mat24_n480_nm16")
    assert s[0] == 60
    o.close()


# ---------------------------------------------------------------------------------------------------
# direct comparison with the reference objects
# ---------------------------------------------------------------------------------------------------
@needs_ref
@pytest.mark.parametrize("rel,kn", [("matrices/Mat26_N48_M16", False), ("matrices/Ahmed_64800_R34_GF16", False),
                                    ("matrices/KN/N96_K48_GF256.txt", True), ("matrices/KN/N576_K480_GF64.txt", True)])
def test_tables_and_graph_equal_reference(rel, kn):
    p = matrix_path(rel)
    r = ol.RefShim(p, kn=kn, n_m=8)
    o = ol.Oracle(p, 2 if kn else 1)
    assert (r.N, r.M, r.GF, r.logGF, r.E, r.K) == (o.N, o.M, o.GF, o.logGF, o.E, o.K)
    assert r.rate == o.rate
    assert (r.row_deg == o.row_deg).all() and (r.col == o.col).all() and (r.val == o.val).all()
    assert (r.bingf == o.bingf).all() and (r.addgf == o.addgf).all()
    assert (r.mulgf == o.mulgf).all() and (r.divgf == o.divgf).all()
    o.close()


def _random_lists(rng, B, n_m, GF, short_frac=0.3, tie_frac=0.3):
    """Sorted n_m-lists as ElementaryStep sees them: ascending LLRs starting at 0, distinct symbols, optional
    absent tail (1e5, -1), plus deliberately tied values."""
    llr = np.zeros((B, n_m), np.float32)
    gf = np.zeros((B, n_m), np.int32)
    for b in range(B):
        step = rng.choice([0.25, 0.5, 1.0]) if rng.random() < tie_frac else None
        v = np.cumsum(rng.random(n_m).astype(np.float32) * 2)
        if step:
            v = np.round(v / step) * step
        v = np.sort(v - v[0]).astype(np.float32)
        llr[b] = v
        gf[b] = rng.permutation(GF)[:n_m] if GF >= n_m else rng.integers(0, GF, n_m)
        if rng.random() < short_frac:
            L = int(rng.integers(1, n_m))
            llr[b, L:] = 1e5
            gf[b, L:] = -1
    return llr, gf


@needs_ref
@pytest.mark.parametrize("rel,n_m,nb_oper", [("matrices/N96_K48_GF64", 20, 25), ("matrices/N96_K48_GF64", 8, 6),
                                             ("matrices/Mat26_N48_M16", 5, 40), ("matrices/Ahmed_64800_R34_GF16", 16, 25),
                                             ("matrices/KN/N96_K48_GF256.txt", 32, 70)])
def test_elementary_step_equals_reference(rel, n_m, nb_oper):
    p = matrix_path(rel)
    kn = rel.endswith(".txt")
    r = ol.RefShim(p, kn=kn, n_m=n_m)
    o = ol.Oracle(p, 2 if kn else 1)
    rng = np.random.default_rng(1234)
    B = 1500
    a, ia = _random_lists(rng, B, n_m, o.GF)
    b, ib = _random_lists(rng, B, n_m, o.GF)
    for k in range(B):
        x = r.elementary_step(a[k], b[k], ia[k], ib[k], n_m, nb_oper)
        y = o.elementary_step(a[k], b[k], ia[k], ib[k], n_m, nb_oper)
        assert x[0].tobytes() == y[0].tobytes() and (x[1] == y[1]).all(), k
    o.close()


@needs_ref
@pytest.mark.parametrize("rel,n_m,nb_oper,offset", [("matrices/N96_K48_GF64", 20, 25, 0.3), ("matrices/Mat28_N72_M18", 12, 25, 1.0),
                                                    ("matrices/Mat212_N96_M16", 16, 20, 0.0),
                                                    ("matrices/Ahmed_64800_R34_GF16", 16, 25, 0.3),
                                                    ("matrices/KN/N96_K48_GF256.txt", 20, 25, 0.3)])
def test_check_node_equals_reference(rel, n_m, nb_oper, offset):
    p = matrix_path(rel)
    kn = rel.endswith(".txt")
    r = ol.RefShim(p, kn=kn, n_m=n_m)
    o = ol.Oracle(p, 2 if kn else 1)
    rng = np.random.default_rng(99)
    dc = int(o.row_deg[0])
    for k in range(300):
        node = int(rng.integers(0, o.M))
        vl, vg = _random_lists(rng, dc, n_m, o.GF, short_frac=0.0)
        x = r.check_node(node, vl, vg, nb_oper, offset)
        y = o.check_node(node, vl, vg, n_m, nb_oper, offset)
        assert x[0].tobytes() == y[0].tobytes() and (x[1] == y[1]).all(), k
    o.close()


def _synd_lists(rng, dc, n_m, GF, mode):
    """V->C lists as the decode loop produces them: ascending from 0, distinct symbols; mode 1 adds exact ties."""
    vl = np.sort(rng.random((dc, n_m)).astype(np.float32) * np.float32(rng.choice([1, 5, 20])), axis=1)
    if mode == 1:
        vl = (np.round(vl * 4) / 4).astype(np.float32)
    vl[:, 0] = 0
    vl = np.sort(vl, axis=1).astype(np.float32)
    vg = np.stack([rng.permutation(GF)[:n_m] for _ in range(dc)]).astype(np.int32)
    return vl, vg


@needs_ref
@pytest.mark.parametrize("rel,n_m,dd,trunc,n_cv", [("matrices/Mat24_N480_M240", 16, (15, 10, 5), 0, 20),
                                                   ("matrices/AD_64800_R12_GF256", 20, (19, 15, 5), 1000, 25),
                                                   ("matrices/N96_K48_GF64", 20, (19, 15, 5), 300, 25),
                                                   ("matrices/Mat24_N48_M24", 8, (7, 4, 2), 0, 6)])
def test_syndrome_check_node_equals_reference(rel, n_m, dd, trunc, n_cv):
    """build_config_table + sort_config_table + syndrome_ems (presorting_mvc, sorting, bayes) vs the compiled reference."""
    path = matrix_path(rel)
    o = ol.Oracle(path, 1)
    r = ol.RefShim(path, False, n_m)
    dc = int(o.row_deg[0])
    cfg = o.build_config_table(dc, *dd, trunc)
    ref_cfg = r.build_config(dc, *dd, trunc)
    assert cfg.shape == ref_cfg.shape and (cfg == ref_cfg).all()
    rng = np.random.default_rng(11)
    for it in range(120):
        node = int(rng.integers(0, o.M))
        vl, vg = _synd_lists(rng, dc, n_m, o.GF, it % 3)
        cl, cg = o.check_node_syndrome(node, vl, vg, n_m, cfg, 0.3, n_cv)
        rl, rg = r.syndrome_ems(node, vl, vg, dc, 0.3, n_cv)
        assert cl.tobytes() == rl.tobytes() and (cg == rg).all()
    o.close()


@needs_ref
def test_frame_source_and_channel_equal_reference():
    p = matrix_path("matrices/Mat24_N96_M48")
    r = ol.RefShim(p, n_m=16, encoder=True)
    o = ol.Oracle(p)
    o.prepare_encoder(); o.rng_default(); r.seed_default()
    for _ in range(5):
        cw_r, nb_r = r.random_codeword()
        il_r, ig_r = r.channel(nb_r, 2.25)
        cw_o, nb_o = o.random_codeword()
        noisy = o.channel_noise(nb_o, 2.25)
        llr = o.channel_llr(noisy, o.sigma(2.25))
        il_o, ig_o = o.sort_intrinsic(llr)
        assert (cw_r == cw_o).all() and (nb_r == nb_o).all()
        assert il_r.tobytes() == il_o.tobytes() and (ig_r == ig_o).all()
        d_r, s_r = r.decision_syndrome(llr)
        assert (d_r == o.decision(llr)).all() and s_r == o.syndrome(d_r)
        assert o.syndrome(cw_o) == 0
    o.close()


def test_apsk64_channel_port_equals_committed_reference_vectors():
    """64-APSK channel (ModelChannel_AWGN_64, channel.c:112-312): the oracle port on the unseeded drand48
    stream against tests/golden/channels/apsk64_n96_gf64.npz (sorted intrinsic LLR / GF recorded from the reference)."""
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "channels", "apsk64_n96_gf64.npz"))
    o = ol.Oracle(matrix_path(str(z["matrix"])))
    o.prepare_encoder(); o.rng_default()
    ebn = float(z["ebn"])
    for f in range(z["nbin"].shape[0]):
        _, nb = o.random_codeword()
        assert (nb == z["nbin"][f]).all()
        noisy = o.channel_noise_apsk64(nb, ebn)
        il, ig = o.sort_intrinsic(o.channel_llr_apsk64(noisy, o.sigma_apsk64(ebn)))
        assert il.tobytes() == z["illr"][f].tobytes() and (ig == z["igf"][f]).all(), f
    mod = o.apsk64_table()
    assert abs(float((mod.astype(np.float64) ** 2).sum()) / 64 - 1.0) < 1e-6          # average power 1, channel.c:205-211
    o.close()


@needs_ref
def test_apsk64_channel_port_equals_reference():
    p = matrix_path("matrices/N96_K48_GF64")
    r = ol.RefShim(p, n_m=20, encoder=True)
    o = ol.Oracle(p)
    o.prepare_encoder(); o.rng_default(); r.seed_default()
    for ebn in (5.0, 9.0, 14.0):
        for _ in range(8):
            _, nb_r = r.random_codeword()
            il_r, ig_r = r.channel_apsk64(nb_r, ebn)
            _, nb_o = o.random_codeword()
            noisy = o.channel_noise_apsk64(nb_o, ebn)
            il_o, ig_o = o.sort_intrinsic(o.channel_llr_apsk64(noisy, o.sigma_apsk64(ebn)))
            assert (nb_r == nb_o).all() and il_r.tobytes() == il_o.tobytes() and (ig_r == ig_o).all()
    o.close()


def test_rng_skip_is_jump_ahead():
    o = ol.Oracle(matrix_path("matrices/N96_K48_GF64"))
    o.rng_default()
    seq = [o.drand48() for _ in range(1000)]
    for n in (0, 1, 7, 999):
        o.rng_default(); o.rng_skip(n)
        assert o.drand48() == seq[n]
    o.close()
