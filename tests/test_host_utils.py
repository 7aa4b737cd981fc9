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


def test_halton_rr2_against_scipy_and_the_published_permutations():
    """Pin for the sample generator's quasi-random sequence (generate_dla_samples.m:8-13, MATLAB scramble(haltonset, 'rr2')):
    (1) the RR2 digit permutations equal the tables Kocis & Whiten (1997, ACM TOMS 23, Table IV) print for the first primes;
    (2) the index -> digits part equals SciPy's independent Halton implementation: un-permuting our digits must give
    scipy.stats.qmc.Halton(scramble=False), whose first point is the origin like haltonset's."""
    from scipy.stats import qmc
    from gp_dla_detection_b200 import dla_samples as D
    published = {2: [0, 1], 3: [0, 2, 1], 5: [0, 4, 2, 1, 3], 7: [0, 4, 2, 6, 1, 5, 3], 11: [0, 8, 4, 2, 10, 6, 1, 9, 5, 3, 7],
                 13: [0, 8, 4, 12, 2, 10, 6, 1, 9, 5, 3, 11, 7]}
    for b, perm in published.items():
        assert D.rr2_permutation(b).tolist() == perm, b
    n, dims = 5000, 3
    ours = D.halton_rr2(n, dims)
    ref = qmc.Halton(d=dims, scramble=False).random(n)
    assert np.array_equal(ref[0], np.zeros(dims))
    for d, b in enumerate((2, 3, 5)):
        m = int(np.ceil(np.log(n) / np.log(b))) + 1            # digits that can be non-zero for indices < n
        N = np.rint(ref[:, d] * float(b) ** m).astype(np.int64)   # SciPy's radical inverse as an exact integer: digits reversed
        assert np.max(np.abs(N / float(b) ** m - ref[:, d])) < 1e-15
        perm = D.rr2_permutation(b)
        expect = np.zeros(n)
        for i in range(m):                                      # digit i (weight b^-(i+1)) of SciPy's point, permuted
            dig = (N // b ** (m - 1 - i)) % b
            expect += perm[dig] * float(b) ** -(i + 1)
        assert np.max(np.abs(ours[:, d] - expect)) < 1e-15, (b, np.max(np.abs(ours[:, d] - expect)))
    # skip: a later start index is the same sequence shifted
    assert np.array_equal(D.halton_rr2(100, 2, skip=1000), D.halton_rr2(1100, 2)[1000:])


def test_processed_mat_has_the_reference_variables_and_layout(tmp_path):
    """write_processed_mat: the variable names of process_qsos.m:236-244 / ...meanflux.m:498-510 in MATLAB's shapes; read
    back through scipy.io.loadmat and, transposed as h5py presents a v7.3 file, indexed the way qso_loader.py:84-110 does."""
    from scipy.io import loadmat
    from gp_dla_detection_b200 import catalog_writer as W
    rng = np.random.default_rng(0)
    Q, S, MD = 5, 7, 3
    single = {n: rng.standard_normal(Q) for n in ("min_z_dlas", "max_z_dlas", "log_priors_no_dla", "log_priors_dla",
              "log_likelihoods_no_dla", "log_likelihoods_dla", "log_posteriors_no_dla", "log_posteriors_dla", "p_no_dlas", "p_dlas")}
    single["model_posteriors"] = rng.random((Q, 2)); single["sample_log_likelihoods_dla"] = rng.standard_normal((Q, S))
    single["map_inds"] = np.arange(Q)                                             # not a saved variable: must be left out
    info = dict(training_release="dr12q", test_set_name="dr12q", num_lines=3, max_z_cut=0.01, test_ind=np.ones((Q, 1), dtype=bool))
    path = str(tmp_path / "processed_qsos_dr12q.mat")
    W.write_processed_mat(path, single, info)
    m = loadmat(path)
    assert set(W.SINGLE_VARIABLES) <= set(m) and "map_inds" not in m
    assert m["p_dlas"].shape == (Q, 1) and m["model_posteriors"].shape == (Q, 2) and m["sample_log_likelihoods_dla"].shape == (Q, S)
    assert m["training_release"][0] == "dr12q" and int(m["num_lines"][0, 0]) == 3
    h5 = W.h5py_view({k: v for k, v in m.items() if not k.startswith("__")})
    assert np.array_equal(h5["p_dlas"][0, :], single["p_dlas"])                   # qso_loader.py:88
    assert np.array_equal(h5["model_posteriors"].T, single["model_posteriors"])   # qso_loader.py:86
    assert h5["sample_log_likelihoods_dla"].shape == (S, Q)
    # multi-DLA run: level axes, 1-based base_sample_inds with 0 for levels never reached, all_exceptions
    multi = {n: rng.standard_normal(Q) for n in ("min_z_dlas", "max_z_dlas", "log_priors_no_dla", "log_priors_lls",
             "log_likelihoods_no_dla", "log_likelihoods_lls", "log_posteriors_no_dla", "log_posteriors_lls", "p_no_dlas", "p_dlas", "p_lls")}
    for n in ("log_priors_dla", "log_likelihoods_dla", "log_posteriors_dla"):
        multi[n] = rng.standard_normal((Q, MD))
    multi["log_likelihoods_dla"][1, 1:] = np.nan                                  # quasar 1 stopped after level 1
    multi["min_z_dlas"][4] = np.nan                                               # quasar 4: no usable pixel
    multi["log_likelihoods_dla"][4] = np.nan
    multi["model_posteriors"] = rng.random((Q, MD + 2))
    multi["MAP_z_dlas"] = rng.random((Q, MD, MD)); multi["MAP_log_nhis"] = rng.random((Q, MD, MD))
    multi["sample_log_likelihoods_dla"] = rng.standard_normal((Q, S, MD)); multi["sample_log_likelihoods_lls"] = rng.standard_normal((Q, S))
    multi["base_sample_inds"] = rng.integers(0, S, (Q, S, MD - 1)).astype(np.int32)
    path2 = str(tmp_path / "processed_qsos_multi_meanflux.mat")
    W.write_processed_mat(path2, multi, multi=True)
    m2 = loadmat(path2)
    assert set(W.MULTI_VARIABLES) <= set(m2)
    b = m2["base_sample_inds"]
    assert b.dtype == np.uint32 and b.shape == (Q, S, MD - 1)
    assert np.array_equal(b[0], multi["base_sample_inds"][0] + 1)                 # 1-based
    assert np.all(b[1, :, 0] == multi["base_sample_inds"][1, :, 0] + 1) and np.all(b[1, :, 1] == 0) and np.all(b[4] == 0)
    assert np.isnan(m2["all_exceptions"][:4]).all() and m2["all_exceptions"][4, 0] == 1
    h5 = W.h5py_view({k: v for k, v in m2.items() if not k.startswith("__")})
    assert h5["MAP_log_nhis"].T.shape == (Q, MD, MD) and h5["sample_log_likelihoods_dla"].shape == (MD, S, Q)   # qso_loader.py:96-97
    small = W.matlab_arrays(multi, multi=True, small_file=True)
    assert "sample_log_likelihoods_dla" not in small and "base_sample_inds" not in small and "p_lls" in small


def test_ascii_results_print_nan_like_matlab(tmp_path):
    """A quasar without usable pixels has NaN results (process_qsos.m:74-82); MATLAB's fprintf prints them as NaN padded
    to the field width, not as C's nan."""
    from gp_dla_detection_b200 import catalog_writer as W
    res = {k: np.array([2.0, np.nan]) for k in ("min_z_dlas", "max_z_dlas", "log_priors_no_dla", "log_priors_dla",
           "log_likelihoods_no_dla", "log_likelihoods_dla", "p_dlas", "map_z_dlas", "map_log_nhis")}
    res["model_posteriors"] = np.array([[0.25, 0.75], [np.nan, np.nan]])
    path = tmp_path / "r.dat"
    W.write_results(str(path), res, [11, 22])
    lines = path.read_text().splitlines()
    assert lines[0] == "000000011 2.0000 2.0000  2.00000  2.00000  2.00000e+00  2.00000e+00 2.50000e-001 7.50000e-001 2.0000 02.0000"
    assert lines[1] == "000000022    NaN    NaN      NaN      NaN          NaN          NaN NaN NaN    NaN     NaN"
