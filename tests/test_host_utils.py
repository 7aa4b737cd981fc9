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
