"""GP training objective (objective.m + spectrum_loss.m): oracle pinning on CPU, CUDA parity on the GPU."""
import numpy as np
import pytest


def test_oracle_matches_dense_density_and_finite_differences():
    from scipy.stats import multivariate_normal
    from oracle import objective_oracle as OB
    x, y, lya, nv = OB.make_training_set(3, num_pixels=60, k=4, seed=1, missing=0.2)
    x[-3:] += [0.2, 0.1, -0.05]                # off the prior means
    P, k = 60, 4
    f, g = OB.objective(x, y, lya, nv, priors=False)
    M = x[:P * k].reshape(k, P).T
    om2 = np.exp(2 * x[P * k:P * (k + 1)]); c0, t0, b = np.exp(x[-3:])
    ref = 0.0
    for i in range(3):
        ind = ~np.isnan(y[i])
        d = nv[i, ind] + om2[ind] * (1 - np.exp(-t0 * lya[i, ind] ** b) + c0) ** 2
        ref -= multivariate_normal.logpdf(y[i, ind], np.zeros(ind.sum()), M[ind] @ M[ind].T + np.diag(d))
    assert abs(f - ref) < 1e-9 * abs(ref)
    rng = np.random.default_rng(0)
    for j in list(rng.integers(0, x.size - 3, 12)) + [x.size - 3, x.size - 2, x.size - 1]:
        h = 1e-6 * max(1.0, abs(x[j]))
        xp, xm = x.copy(), x.copy(); xp[j] += h; xm[j] -= h
        fd = (OB.objective(xp, y, lya, nv, priors=False)[0] - OB.objective(xm, y, lya, nv, priors=False)[0]) / (2 * h)
        assert abs(fd - g[j]) < 1e-5 * max(1.0, abs(g[j])), (j, fd, g[j])
    # the priors only touch the tau_0 and beta entries
    gp = OB.objective(x, y, lya, nv)[1]
    assert np.array_equal(gp[:-2], g[:-2]) and gp[-2] != g[-2] and gp[-1] != g[-1]


@pytest.mark.gpu
@pytest.mark.parametrize("k,P,N", [(20, 1217, 40), (10, 333, 17), (40, 200, 5)])
def test_objective_matches_oracle(k, P, N):
    from gp_dla_detection_b200 import api
    from oracle import objective_oracle as OB
    x, y, lya, nv = OB.make_training_set(N, num_pixels=P, k=k, seed=k)
    y[N // 2, :] = np.nan                      # a spectrum without any observed pixel contributes nothing
    x[-3:] += [0.2, 0.1, -0.05]                # move off the prior means so every gradient block is exercised
    f, g = api.objective(x, y, lya, nv)
    fr, gr = OB.objective(x, y, lya, nv)
    assert abs(f - fr) < 1e-11 * abs(fr)
    scale = np.abs(gr).max()
    assert np.max(np.abs(g - gr)) < 1e-10 * scale, np.max(np.abs(g - gr)) / scale
    assert np.allclose(g[-3:], gr[-3:], rtol=1e-9)
    # device-resident evaluator: same numbers, data uploaded once
    ev = api.TrainingObjective(y, lya, nv, k)
    f2, g2 = ev(x)
    assert abs(f2 - f) < 1e-12 * abs(f) and np.max(np.abs(g2 - g)) < 1e-12 * scale
    ev.close()



def forest_constants():
    """all_transition_wavelengths (Angstrom) and all_oscillator_strengths of multi_dlas/set_parameters_multi.m:76-142 (the
    same Lyman-series data as voigt.c:31-99, which the oracle tabulates)."""
    from oracle import process_qsos_oracle as O
    return O.TRANSITION_WAVELENGTHS * 1e8, O.OSCILLATOR_STRENGTHS


def test_lyseries_oracle_matches_dense_density_and_finite_differences():
    """The Lyman-series objective: value against the dense multivariate normal with the optical depth written out line
    by line, tau_0 gradient by finite differences (the reference's beta gradient uses log(1 + z_Lya) for every series
    member, spectrum_loss_lyseries.m:89, and is NOT the derivative of f: it is restated as written)."""
    from scipy.stats import multivariate_normal
    from oracle import objective_oracle as OB
    tw, osc = forest_constants()
    x, y, lya, nv = OB.make_training_set(3, num_pixels=60, k=4, seed=2, missing=0.2)
    lya[:, -1] = np.where(np.isnan(lya[:, -1]), 1 + 2.9, lya[:, -1])     # 1 + z_qso comes from the last column
    x[-3:] += [0.2, 0.1, -0.05]
    P, k, NL = 60, 4, 31
    f, g = OB.objective(x, y, lya, nv, priors=False, lyseries=(NL, tw, osc))
    f0 = OB.objective(x, y, lya, nv, priors=False)[0]
    assert abs(f - f0) > 1e-6 * abs(f0)                                  # the series members do change the model
    M = x[:P * k].reshape(k, P).T
    om2 = np.exp(2 * x[P * k:P * (k + 1)]); c0, t0, b = np.exp(x[-3:])
    ref = 0.0
    for i in range(3):
        ind = ~np.isnan(y[i])
        z1 = lya[i, ind]
        od = t0 * z1 ** b
        for l in range(1, NL):
            zl = tw[0] * z1 / tw[l]
            od = od + np.where(zl <= lya[i, -1], t0 * tw[l] * osc[l] / (tw[0] * osc[0]) * zl ** b, 0.0)
        d = nv[i, ind] + om2[ind] * (1 - np.exp(-od) + c0) ** 2
        ref -= multivariate_normal.logpdf(y[i, ind], np.zeros(ind.sum()), M[ind] @ M[ind].T + np.diag(d))
    assert abs(f - ref) < 1e-9 * abs(ref)
    rng = np.random.default_rng(0)
    for j in list(rng.integers(0, x.size - 3, 10)) + [x.size - 3, x.size - 2]:      # M, log omega, log c_0, log tau_0
        h = 1e-6 * max(1.0, abs(x[j]))
        xp, xm = x.copy(), x.copy(); xp[j] += h; xm[j] -= h
        fd = (OB.objective(xp, y, lya, nv, priors=False, lyseries=(NL, tw, osc))[0]
              - OB.objective(xm, y, lya, nv, priors=False, lyseries=(NL, tw, osc))[0]) / (2 * h)
        assert abs(fd - g[j]) < 1e-5 * max(1.0, abs(g[j])), (j, fd, g[j])
    # one line = the base objective, bit for bit
    f1, g1 = OB.objective(x, y, lya, nv, lyseries=(1, tw, osc))
    fb, gb = OB.objective(x, y, lya, nv)
    assert f1 == fb and np.array_equal(g1, gb)


@pytest.mark.gpu
@pytest.mark.parametrize("k,P,N,NL", [(20, 1217, 40, 31), (10, 333, 17, 5), (40, 200, 5, 31)])
def test_objective_lyseries_matches_oracle(k, P, N, NL):
    from gp_dla_detection_b200 import api
    from oracle import objective_oracle as OB
    tw, osc = forest_constants()
    x, y, lya, nv = OB.make_training_set(N, num_pixels=P, k=k, seed=100 + k)
    some = np.isnan(lya[:, -1]); some[::3] = False
    lya[:, -1] = np.where(some, lya[:, -1], np.nanmax(lya, axis=1))       # most rows carry 1 + z_qso, some keep NaN there
    y[N // 2, :] = np.nan
    x[-3:] += [0.2, 0.1, -0.05]
    f, g = api.objective_lyseries(x, y, lya, nv, NL, tw, osc)
    fr, gr = OB.objective(x, y, lya, nv, lyseries=(NL, tw, osc))
    assert abs(f - fr) < 1e-11 * abs(fr)
    scale = np.abs(gr).max()
    assert np.max(np.abs(g - gr)) < 1e-10 * scale, np.max(np.abs(g - gr)) / scale
    assert np.allclose(g[-3:], gr[-3:], rtol=1e-9)
    ev = api.TrainingObjective(y, lya, nv, k, num_forest_lines=NL, all_transition_wavelengths=tw, all_oscillator_strengths=osc)
    f2, g2 = ev(x)
    f3, g3 = ev(x)
    assert f2 == f and np.array_equal(g2, g)           # two-stage reduction: the same bits on every evaluation
    assert f3 == f2 and np.array_equal(g3, g2)
    ev.close()
