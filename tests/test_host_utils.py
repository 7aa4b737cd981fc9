"""CPU tests of the host-side helpers either side of the hot path: DLA sample
generation (generate_dla_samples.m) -- parity unpinned against MATLAB's toolbox, so these are property tests."""
import numpy as np

from gp_dla_detection_b200 import dla_samples as D


def test_rr2_permutations_and_halton_uniformity():
    for b in (2, 3, 5, 7):
        p = D.rr2_permutation(b)
        assert sorted(p) == list(range(b)) and p[0] == 0
    h = D.halton_rr2(4096, 3)
    assert h.shape == (4096, 3) and h.min() >= 0.0 and h.max() < 1.0
    assert np.array_equal(h[0], np.zeros(3))                      # haltonset starts at the origin
    for d in range(3):                                            # low discrepancy: every 1/16 bin equally filled
        counts = np.histogram(h[:, d], bins=16, range=(0, 1))[0]
        assert counts.max() - counts.min() <= 16
    assert abs(np.corrcoef(h[:, 0], h[:, 1])[0, 1]) < 0.02


def test_generated_samples_follow_the_mixture_prior():
    f = np.array([-0.03, -0.93, 30.4])                            # a quadratic log-pdf like the fit of :33-34
    s = D.generate_dla_samples(3000, log_pdf_poly=f, alpha=0.9)
    x = s["log_nhi_samples"]
    assert x.min() >= 20.0 and x.max() <= 25.0 and np.allclose(s["nhi_samples"], 10.0 ** x)
    # empirical CDF of the quasi-random draws against the analytic mixture CDF
    grid = np.linspace(20, 25, 20001)
    pdf_fit = np.exp(np.polyval(f, grid)); pdf_fit /= np.trapezoid(pdf_fit, grid)
    pdf = 0.9 * pdf_fit + 0.1 * ((grid <= 23.0) / 3.0)
    cdf = np.concatenate([[0], np.cumsum(0.5 * (pdf[1:] + pdf[:-1]) * np.diff(grid))])
    emp = np.searchsorted(np.sort(x), grid, side="right") / x.size
    assert np.max(np.abs(emp - cdf)) < 2e-3                      # quasi-Monte-Carlo: far below 1/sqrt(n)
    assert np.array_equal(s["offset_samples"], D.halton_rr2(3000, 2)[:, 0])


def test_lls_samples_and_partition_functions():
    """multi_dlas/set_lls_parameters.m: sub-DLA samples uniform on [19.5, 20) from Halton dimension 3; Z_lls + Z_dla
    is the mixture's mass on [19.5, 23], which misses only the fitted density's tail above 23."""
    f = np.array([-0.03, -0.93, 30.4])
    s = D.generate_lls_samples(2048, log_pdf_poly=f)
    x = s["lls_log_nhi_samples"]
    assert x.shape == (2048,) and x.min() >= 19.5 and x.max() < 20.0
    assert np.allclose(s["lls_nhi_samples"], 10.0 ** x)
    assert np.array_equal(s["lls_offset_samples"], D.halton_rr2(2048, 3)[:, 2])
    assert np.array_equal(s["offset_samples"], D.halton_rr2(2048, 3)[:, 0])
    counts = np.histogram(x, bins=8, range=(19.5, 20.0))[0]
    assert counts.max() - counts.min() <= 8                       # uniform in log N
    # independent quadrature of the same mixture
    grid = np.linspace(19.5, 25.0, 550001)
    peak = 20.03269
    un = np.where(grid > peak, np.exp(np.polyval(f, grid)), np.exp(np.polyval(f, peak)))
    Z = np.trapezoid(un, grid)
    pdf = 0.97 * un / Z + 0.03 * ((grid <= 23.0) / 3.5)
    z_lls = np.trapezoid(pdf[grid <= 20.0], grid[grid <= 20.0])
    sel = (grid >= 20.0) & (grid <= 23.0)
    z_dla = np.trapezoid(pdf[sel], grid[sel])
    assert abs(s["Z_lls"] - z_lls) < 1e-5 and abs(s["Z_dla"] - z_dla) < 1e-5
    assert 0.0 < s["Z_lls"] < s["Z_dla"] < 1.0 and s["Z_lls"] + s["Z_dla"] < 1.0


def test_json_catalogues(tmp_path):
    """qso_loader.py:1927-2090: record layout, num_dlas from the arg-max model, sub-DLA / null folded into p_no_dla."""
    import json
    from gp_dla_detection_b200 import catalog_writer as cw
    mp = np.array([[0.7, 0.1, 0.15, 0.05], [0.1, 0.6, 0.2, 0.1], [0.05, 0.05, 0.2, 0.7], [0.1, 0.1, 0.6, 0.2]])
    MAPz = np.arange(4 * 2 * 2, dtype=float).reshape(4, 2, 2) / 10 + 2.0
    MAPn = np.arange(4 * 2 * 2, dtype=float).reshape(4, 2, 2) / 10 + 20.0
    res = dict(model_posteriors=mp, p_no_dlas=mp[:, 0] + mp[:, 1], p_dlas=1 - mp[:, 0] - mp[:, 1],
               min_z_dlas=np.full(4, 2.0), max_z_dlas=np.full(4, 3.0), MAP_z_dlas=MAPz, MAP_log_nhis=MAPn)
    info = dict(thing_ids=np.array([11, 22, 33, 44]), z_qsos=np.full(4, 3.1), snrs=np.full(4, 5.0),
                ras=np.arange(4.0), decs=-np.arange(4.0), plates=np.array([1, 2, 3, 4]), mjds=np.array([5, 6, 7, 8]),
                fiber_ids=np.array([9, 10, 11, 12]))
    out = cw.write_json_catalogue(str(tmp_path / "p.json"), res, info)
    back = json.load(open(tmp_path / "p.json"))
    assert back == out and [r["num_dlas"] for r in back] == [0, 0, 2, 1]
    assert back[0]["max_model_posterior"] == 0.8 or abs(back[0]["max_model_posterior"] - 0.8) < 1e-15
    assert abs(back[1]["max_model_posterior"] - 0.7) < 1e-15 and back[2]["max_model_posterior"] == 0.7
    assert back[2]["dlas"] == [{"log_nhi": MAPn[2, 1, 0], "z_dla": MAPz[2, 1, 0]},
                               {"log_nhi": MAPn[2, 1, 1], "z_dla": MAPz[2, 1, 1]}]
    assert back[3]["dlas"] == [{"log_nhi": MAPn[3, 0, 0], "z_dla": MAPz[3, 0, 0]}] and back[0]["dlas"] == []
    assert set(back[0]) == {"p_dla", "p_no_dla", "max_model_posterior", "num_dlas", "dlas", "min_z_dla", "max_z_dla",
                            "ra", "snr", "dec", "plate", "mjd", "fiber_id", "thing_id", "z_qso"}
    sub = cw.write_sub_dla_catalogue(str(tmp_path / "s.json"), res, info)
    assert [r["thing_id"] for r in sub] == [22] and sub[0]["p_sub_dla"] == 0.6
    assert json.load(open(tmp_path / "s.json")) == sub
