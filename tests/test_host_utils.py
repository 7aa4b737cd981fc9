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
