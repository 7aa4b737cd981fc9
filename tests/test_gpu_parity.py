"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the golden vectors.

Tolerances are BASELINE.json's north_star: same MAP sample index, log-likelihoods within 1e-8
relative, p_dla within 1e-6 absolute (the asserts below are tighter where the path allows)."""
import os

import numpy as np
import pytest

from conftest import boss_grid

LL_RTOL = 1e-8        # north_star tolerance on log-likelihoods
LL_RTOL_TIGHT = 1e-10  # what this implementation is held to
P_ATOL = 1e-6         # north_star tolerance on p_dla
PROFILE_ATOL = 5e-15  # absorption profile, absolute (values in [0, 1])


def load_golden_problem(golden_dir):
    g = np.load(os.path.join(golden_dir, "process_qsos_small.npz"))
    model = dict(rest_wavelengths=g["model_rest_wavelengths"], mu=g["model_mu"], M=g["model_M"],
                 log_omega=g["model_log_omega"], log_c_0=g["model_scalars"][0], log_tau_0=g["model_scalars"][1],
                 log_beta=g["model_scalars"][2])
    samples = {k: g[k] for k in ("offset_samples", "log_nhi_samples", "nhi_samples")}
    prior = dict(z_qsos=g["prior_z_qsos"], dla_ind=g["prior_dla_ind"])
    spectra = {k: [g["%s_%d" % (k, q)] for q in range(3)]
               for k in ("all_wavelengths", "all_flux", "all_noise_variance", "all_pixel_mask")}
    spectra["z_qsos"] = g["z_qsos"]
    expect = {k[4:]: g[k] for k in g.files if k.startswith("out_")}
    return model, samples, spectra, prior, expect


def assert_parity(res, ref, check_samples=True):
    assert np.array_equal(res["map_inds"], ref["map_inds"])
    for k in ("log_likelihoods_no_dla", "log_likelihoods_dla", "log_posteriors_no_dla", "log_posteriors_dla"):
        assert np.allclose(res[k], ref[k], rtol=LL_RTOL_TIGHT, atol=0, equal_nan=True), k
    if check_samples:
        a, b = res["sample_log_likelihoods_dla"], ref["sample_log_likelihoods_dla"]
        assert np.allclose(a, b, rtol=LL_RTOL_TIGHT, atol=0, equal_nan=True)
    for k in ("p_dlas", "p_no_dlas"):
        assert np.allclose(res[k], ref[k], rtol=0, atol=P_ATOL, equal_nan=True), k
    assert np.allclose(res["model_posteriors"], ref["model_posteriors"], rtol=0, atol=P_ATOL, equal_nan=True)
    for k in ("min_z_dlas", "max_z_dlas", "log_priors_no_dla", "log_priors_dla", "map_z_dlas", "map_log_nhis"):
        assert np.allclose(res[k], ref[k], rtol=1e-14, atol=0, equal_nan=True), k


pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from gp_dla_detection_b200 import api as A
    return A


def test_voigt_against_oracle(api):
    from oracle import process_qsos_oracle as O
    lam = boss_grid()
    rng = np.random.default_rng(3)
    cases = [(2.3, 1e21, 3), (2.1, 10 ** 20.3, 31), (2.5, 1e23, 3), (2.9, 1e20, 1), (2.2, 10 ** 19.5, 3)]
    cases += [(rng.uniform(1.95, 2.95), 10 ** rng.uniform(20, 23), int(rng.choice([1, 2, 3, 7, 31]))) for _ in range(25)]
    for z, N, nl in cases:
        a, b = api.voigt(lam, z, N, nl), O.voigt(lam, z, N, nl)
        assert a.shape == (lam.size - 6,)
        assert np.max(np.abs(a - b)) <= PROFILE_ATOL, (z, N, nl)
    assert np.array_equal(api.voigt(lam, 2.2, 0.0, 3), np.ones(lam.size - 6))   # N = 0: exactly no absorption
    assert api.voigt(lam[:7], 2.2, 1e20, 3).shape == (1,)                       # minimum length


def test_voigt_against_reference_golden_vectors(api, golden_dir):
    g = np.load(os.path.join(golden_dir, "voigt_reference.npz"))
    for i, (z, N, nl) in enumerate(g["cases"]):
        lam = g["lambdas"] if z < 3.5 else g["lambdas_hi"]
        assert np.max(np.abs(api.voigt(lam, z, N, int(nl)) - g["profile_%d" % i])) <= PROFILE_ATOL


def test_voigt_core_pixels_dense_sweep(api):
    """Sweep the line centre across a pixel in fine steps so every core-table interval is hit."""
    from oracle import process_qsos_oracle as O
    lam = boss_grid(200, 3.60)
    for nl, logn in ((1, 12.5), (3, 13.0), (3, 14.0)):   # unsaturated: the core shape is visible
        for dz in np.linspace(0, 3e-4, 41):
            z = 2.3 + dz
            a, b = api.voigt(lam, z, 10 ** logn, nl), O.voigt(lam, z, 10 ** logn, nl)
            assert np.max(np.abs(a - b)) <= PROFILE_ATOL


def test_voigt_errors(api):
    from gp_dla_detection_b200._lib import GpdlaError
    lam = boss_grid(50)
    for bad in (dict(lambdas=lam[:6], z=2.0, N=1e20, num_lines=3), dict(lambdas=lam, z=2.0, N=1e20, num_lines=0),
                dict(lambdas=lam, z=2.0, N=1e20, num_lines=32), dict(lambdas=lam, z=-1.5, N=1e20, num_lines=3)):
        with pytest.raises(GpdlaError):
            api.voigt(**bad)


def test_voigt_batch_device(api):
    import torch
    from oracle import process_qsos_oracle as O
    lam = boss_grid(700)
    zs, Ns = np.linspace(2.0, 2.6, 9), 10 ** np.linspace(20, 22.5, 9)
    out = api.voigt_batch(torch.from_numpy(lam).cuda(), torch.from_numpy(zs).cuda(), torch.from_numpy(Ns).cuda(), 3)
    out = out.cpu().numpy()
    for i in range(9):
        assert np.max(np.abs(out[i] - O.voigt(lam, zs[i], Ns[i], 3))) <= PROFILE_ATOL


def test_process_qsos_golden_small(api, golden_dir):
    model, samples, spectra, prior, expect = load_golden_problem(golden_dir)
    res = api.process_qsos(model, samples, spectra, prior)
    assert_parity(res, expect)


def test_process_qsos_full_size_against_oracle(api, synthetic_inputs):
    """Reference configuration: 10 000 samples, k = 20, 3 lines, ~1250-pixel spectra."""
    from oracle import process_qsos_oracle as O
    si = synthetic_inputs
    res = api.process_qsos(si["model"], si["samples"], si["spectra"], si["prior"])
    ref = O.process_qsos(si["model"], si["samples"], si["spectra"], si["prior"], engine="c")
    assert_parity(res, ref)
    # north_star's own tolerance, stated explicitly
    rel = np.abs(res["sample_log_likelihoods_dla"] - ref["sample_log_likelihoods_dla"]) / np.abs(ref["sample_log_likelihoods_dla"])
    assert rel.max() < LL_RTOL
    # injected DLAs are recovered: MAP within one sample spacing of the truth
    sp = si["spectra"]
    for q in np.flatnonzero(~np.isnan(sp["truth_z_dla"])):
        assert res["p_dlas"][q] > 0.99
        assert abs(res["map_z_dlas"][q] - sp["truth_z_dla"][q]) < 5e-3
        assert abs(res["map_log_nhis"][q] - sp["truth_log_nhi"][q]) < 1.0   # noisy spectra: loose on N_HI


def test_fixed_shape_1217_pixel_grid(api, synthetic_inputs):
    from gp_dla_detection_b200 import synthetic as syn
    from oracle import process_qsos_oracle as O
    si = synthetic_inputs
    sp = syn.make_spectra(si["model"], 2, seed=11, fixed_shape=True)
    sub = np.arange(0, 10000, 10)
    samples = {k: v[sub] for k, v in si["samples"].items()}
    res = api.process_qsos(si["model"], samples, sp, si["prior"])
    ref = O.process_qsos(si["model"], samples, sp, si["prior"], engine="c")
    assert_parity(res, ref)


def test_more_lyman_lines(api, synthetic_inputs):
    from gp_dla_detection_b200.params import Parameters
    from oracle import process_qsos_oracle as O
    si = synthetic_inputs
    sp = {k: v[:1] for k, v in si["spectra"].items()}
    sub = np.arange(0, 10000, 40)
    samples = {k: v[sub] for k, v in si["samples"].items()}
    for nl in (1, 5, 31):
        res = api.process_qsos(si["model"], samples, sp, si["prior"], params=Parameters(num_lines=nl))
        ref = O.process_qsos(si["model"], samples, sp, si["prior"], num_lines=nl, engine="c")
        assert_parity(res, ref)


def test_ragged_masked_and_empty_spectra(api, synthetic_inputs):
    from oracle import process_qsos_oracle as O
    si = synthetic_inputs
    sp = {k: ([x.copy() for x in v] if isinstance(v, list) else v.copy()) for k, v in si["spectra"].items()}
    for k in ("all_wavelengths", "all_flux", "all_noise_variance", "all_pixel_mask"):
        sp[k][0] = sp[k][0][500:]                      # blue end missing (BOSS coverage limit)
        sp[k][2] = sp[k][2][:900]                      # red end missing
    sp["all_pixel_mask"][1][200:260] = True            # a masked block
    sp["all_flux"][1][200:260] = np.nan                # garbage under the mask must not leak
    sp["all_pixel_mask"][3][:] = True                  # nothing usable: NaN results (process_qsos.m:74-82)
    sub = np.arange(0, 10000, 20)
    samples = {k: v[sub] for k, v in si["samples"].items()}
    res = api.process_qsos(si["model"], samples, sp, si["prior"])
    spo = {k: (v[:3] if isinstance(v, list) else v[:3]) for k, v in sp.items()}
    spo["all_flux"][1] = np.where(sp["all_pixel_mask"][1], 0.0, sp["all_flux"][1])
    ref = O.process_qsos(si["model"], samples, spo, si["prior"], engine="c")
    assert_parity({k: v[:3] for k, v in res.items()}, ref)
    for k in ("log_likelihoods_no_dla", "log_likelihoods_dla", "p_dlas", "map_z_dlas", "min_z_dlas"):
        assert np.isnan(res[k][3]), k
    assert res["map_inds"][3] == -1 and np.all(np.isnan(res["sample_log_likelihoods_dla"][3]))


def test_batching_and_device_entry_agree_with_host_entry(api, synthetic_inputs):
    import torch
    from gp_dla_detection_b200 import synthetic as syn
    si = synthetic_inputs
    sp = syn.make_spectra(si["model"], 7, seed=21)
    sub = np.arange(0, 10000, 50)
    samples = {k: v[sub] for k, v in si["samples"].items()}
    one = api.DLAProcessor(si["model"], samples, si["prior"]).process(sp)
    small = api.DLAProcessor(si["model"], samples, si["prior"], batch_quasars=2).process(sp)
    for k in one:
        assert np.array_equal(one[k], small[k], equal_nan=True), k
    proc = api.DLAProcessor(si["model"], samples, si["prior"], batch_quasars=3)
    pad = api.pad_spectra(sp)
    t = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in pad.items()}
    n0 = proc.launch_count
    dev = proc.process_device(t["wavelengths"], t["flux"], t["noise_variance"], t["pixel_mask"], t["lengths"],
                              t["z_qsos"], return_sample_log_likelihoods=True)
    torch.cuda.synchronize()
    # kernels per batch: prepare, [scales + digit operands + (idle) FP64 fallback operand | FP64 Gram operand], fused
    # log-likelihood [+ (idle) FP64 fallback], evidence
    assert proc.launch_count - n0 in (7 * 3, 4 * 3)   # 3 batches; 7 on the INT8 Gram path, 4 on the FP64 one
    for k in one:
        assert np.array_equal(one[k], dev[k].cpu().numpy(), equal_nan=True), k


def test_zero_column_density_reduces_to_null_model(api, synthetic_inputs):
    """Size-independent property at the full 10 000-sample size: N_HI = 0 for every sample makes each
    sample likelihood equal the null likelihood, so the DLA evidence equals the null evidence."""
    si = synthetic_inputs
    s0 = dict(si["samples"]); s0["nhi_samples"] = np.zeros_like(s0["nhi_samples"])
    res = api.process_qsos(si["model"], s0, si["spectra"], si["prior"])
    for q in range(len(si["spectra"]["z_qsos"])):
        assert np.allclose(res["sample_log_likelihoods_dla"][q], res["log_likelihoods_no_dla"][q], rtol=1e-12, atol=0)
    assert np.allclose(res["log_likelihoods_dla"], res["log_likelihoods_no_dla"], rtol=1e-12, atol=0)


def test_sample_permutation_invariance(api, synthetic_inputs):
    """Permuting the samples permutes the sample likelihoods bit-for-bit and leaves the evidence and MAP
    parameters unchanged (checks tile/offset handling at full size)."""
    si = synthetic_inputs
    sp = {k: v[:2] for k, v in si["spectra"].items()}
    perm = np.random.default_rng(5).permutation(10000)
    sperm = {k: v[perm] for k, v in si["samples"].items()}
    a = api.process_qsos(si["model"], si["samples"], sp, si["prior"])
    b = api.process_qsos(si["model"], sperm, sp, si["prior"])
    assert np.array_equal(a["sample_log_likelihoods_dla"][:, perm], b["sample_log_likelihoods_dla"])
    assert np.allclose(a["log_likelihoods_dla"], b["log_likelihoods_dla"], rtol=1e-13, atol=0)
    assert np.array_equal(a["map_z_dlas"], b["map_z_dlas"]) and np.array_equal(a["map_log_nhis"], b["map_log_nhis"])
    assert np.array_equal(perm[b["map_inds"]], a["map_inds"])


def test_rank_and_sample_count_sweep(api, synthetic_inputs):
    """BASELINE.json configs[4]: low-rank dimensions k = 10 and 40 (k = 20 is covered above), 1e3 / 3e4 samples."""
    from gp_dla_detection_b200 import synthetic as syn
    from oracle import process_qsos_oracle as O
    si = synthetic_inputs
    m10 = syn.make_model(10)
    sp = syn.make_spectra(m10, 2, seed=31, dla_fraction=0.5)
    s1k = syn.make_samples(1000)
    assert_parity(api.process_qsos(m10, s1k, sp, si["prior"]), O.process_qsos(m10, s1k, sp, si["prior"], engine="c"))
    m40 = syn.make_model(40)                                      # column-split kernel + stand-alone Cholesky
    sp40 = syn.make_spectra(m40, 3, seed=41, dla_fraction=0.5)
    sp40["all_pixel_mask"][1][:] = True                            # an empty spectrum inside the batch
    r40 = api.process_qsos(m40, s1k, sp40, si["prior"])
    o40 = O.process_qsos(m40, s1k, sp40, si["prior"], engine="c")
    assert_parity(r40, o40)
    assert np.isnan(r40["log_likelihoods_dla"][1]) and np.all(np.isfinite(r40["log_likelihoods_dla"][[0, 2]]))
    s30k = syn.make_samples(30000)
    sp1 = {k: v[:1] for k, v in si["spectra"].items()}
    assert_parity(api.process_qsos(si["model"], s30k, sp1, si["prior"]),
                  O.process_qsos(si["model"], s30k, sp1, si["prior"], engine="c"))


def test_rank_40_int8_contract_passes(api, synthetic_inputs):
    """k = 40 on the INT8 path (default): the producing kernel covers three column blocks and the projection, the
    contract-only kernel the other eight from the stored digit tiles, cholesky_kernel finishes.  Against the oracle and
    against the FP64 DMMA path at full sample count; a zero-noise-variance quasar takes the FP64 list kernel and the
    same Cholesky launch; sample counts that leave partial 128-row tiles."""
    from gp_dla_detection_b200 import synthetic as syn
    from oracle import process_qsos_oracle as O
    si = synthetic_inputs
    m40 = syn.make_model(40)
    sp = syn.make_spectra(m40, 4, seed=43, dla_fraction=0.5)
    sp["all_pixel_mask"][2][::3] = True
    nv = sp["all_noise_variance"][3]
    w = np.asarray(sp["all_wavelengths"][3]) / (1 + sp["z_qsos"][3])
    used = np.flatnonzero(~np.asarray(sp["all_pixel_mask"][3]) & (w >= 911.75) & (w <= 1215.75))
    nv[used[[-2, -4]]] = 0.0                                      # quasar 3 -> FP64 fallback list (pixels no DLA saturates)
    res = {d: api.process_qsos(m40, si["samples"], sp, si["prior"], gram_digits=d) for d in (6, -1)}
    ref = O.process_qsos(m40, si["samples"], sp, si["prior"], engine="c")
    b = ref["sample_log_likelihoods_dla"]
    for d in (6, -1):
        a = res[d]["sample_log_likelihoods_dla"]
        assert np.max(np.abs(a - b) / np.abs(b)) < 1e-11, (d, np.max(np.abs(a - b) / np.abs(b)))
        assert_parity(res[d], ref)
    a6, af = res[6]["sample_log_likelihoods_dla"], res[-1]["sample_log_likelihoods_dla"]
    assert np.max(np.abs(a6 - af) / np.abs(af)) < 2e-12
    assert np.array_equal(a6[3], af[3])                            # the flagged quasar ran the same FP64 kernels
    for S in (1, 127, 129, 300):
        sub = {k: v[np.arange(S) * 5 + 1] for k, v in si["samples"].items()}
        sp2 = {k: v[:2] for k, v in sp.items()}
        assert_parity(api.process_qsos(m40, sub, sp2, si["prior"]), O.process_qsos(m40, sub, sp2, si["prior"], engine="c"))


def test_tiny_and_odd_sample_counts(api, synthetic_inputs):
    """Sample tiles are 32 wide: S = 1, 5, 33 exercise the partial-tile and null-slot handling."""
    from oracle import process_qsos_oracle as O
    si = synthetic_inputs
    sp = {k: v[:2] for k, v in si["spectra"].items()}
    for S in (1, 5, 33):
        sub = np.arange(S) * 7 + 3
        samples = {k: v[sub] for k, v in si["samples"].items()}
        assert_parity(api.process_qsos(si["model"], samples, sp, si["prior"]),
                      O.process_qsos(si["model"], samples, sp, si["prior"], engine="c"))


def test_two_devices_in_one_process(api, synthetic_inputs):
    """Contexts on different GPUs of one process give identical results (per-device constants/attributes)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    si = synthetic_inputs
    sub = np.arange(0, 10000, 25)
    samples = {k: v[sub] for k, v in si["samples"].items()}
    a = api.process_qsos(si["model"], samples, si["spectra"], si["prior"], device=0)
    b = api.process_qsos(si["model"], samples, si["spectra"], si["prior"], device=1)
    for k in a:
        assert np.array_equal(a[k], b[k], equal_nan=True), k


def test_fuzz_extreme_inputs(api, synthetic_inputs):
    """Randomised spectra outside the synthetic generator's comfort zone: noise variances over 10 decades,
    flux outliers and negative flux, heavy masking, redshifts up to 5.5, short spectra."""
    from gp_dla_detection_b200 import synthetic as syn
    from oracle import process_qsos_oracle as O
    si = synthetic_inputs
    rng = np.random.default_rng(2024)
    sp = syn.make_spectra(si["model"], 16, seed=77, dla_fraction=0.5)
    for q in range(16):
        L = len(sp["all_flux"][q])
        sp["all_noise_variance"][q] = 10.0 ** rng.uniform(-6, 4, L)
        sp["all_flux"][q] = sp["all_flux"][q] + rng.standard_normal(L) * np.sqrt(sp["all_noise_variance"][q])
        out = rng.random(L) < 0.01
        sp["all_flux"][q][out] = rng.uniform(-50, 50, np.count_nonzero(out))
        sp["all_pixel_mask"][q] = rng.random(L) < rng.choice([0.0, 0.05, 0.5])
        if q % 4 == 0:
            lo, n_keep = rng.integers(0, 900), rng.integers(210, 400)
            for k in ("all_wavelengths", "all_flux", "all_noise_variance", "all_pixel_mask"):
                sp[k][q] = sp[k][q][lo:lo + n_keep]
    sp["all_wavelengths"][3] = sp["all_wavelengths"][3] * (1 + 5.5) / (1 + float(sp["z_qsos"][3]))   # same rest frame at z = 5.5
    sp["z_qsos"][3] = 5.5
    sub = np.arange(0, 10000, 33)
    samples = {k: v[sub] for k, v in si["samples"].items()}
    res = api.process_qsos(si["model"], samples, sp, si["prior"])
    ref = O.process_qsos(si["model"], samples, sp, si["prior"], engine="c")
    a, b = res["sample_log_likelihoods_dla"], ref["sample_log_likelihoods_dla"]
    assert np.array_equal(np.isnan(a), np.isnan(b))
    assert np.allclose(a, b, rtol=1e-9, atol=0, equal_nan=True)      # north_star: 1e-8
    assert np.allclose(res["log_likelihoods_no_dla"], ref["log_likelihoods_no_dla"], rtol=1e-9, equal_nan=True)
    assert np.allclose(res["p_dlas"], ref["p_dlas"], rtol=0, atol=P_ATOL, equal_nan=True)
    ok = ~np.isnan(ref["log_likelihoods_dla"])
    tie_free = np.array([np.sum(b[q] >= np.nanmax(b[q]) * (1 + 1e-12 * np.sign(-np.nanmax(b[q])))) == 1 if ok[q] else False for q in range(16)])
    assert np.array_equal(res["map_inds"][tie_free], ref["map_inds"][tie_free])


def test_gram_paths_agree(api, synthetic_inputs):
    """The exact-product INT8 tensor-core Gram (6 and 5 digits) against the FP64 DMMA Gram and the oracle at full
    sample count: 6 digits are FP64-equivalent, 5 digits stay 100x inside the north-star tolerance (1e-8)."""
    from gp_dla_detection_b200 import synthetic as syn
    from oracle import process_qsos_oracle as O
    si = synthetic_inputs
    sp = syn.make_spectra(si["model"], 3, seed=5, dla_fraction=0.67)
    res = {d: api.process_qsos(si["model"], si["samples"], sp, si["prior"], gram_digits=d) for d in (-1, 6, 5)}
    ref = O.process_qsos(si["model"], si["samples"], sp, si["prior"], engine="c")
    b = ref["sample_log_likelihoods_dla"]
    for d, tol in ((-1, 1e-11), (6, 1e-11), (5, 1e-9)):
        a = res[d]["sample_log_likelihoods_dla"]
        assert np.max(np.abs(a - b) / np.abs(b)) < tol, (d, np.max(np.abs(a - b) / np.abs(b)))
        assert np.allclose(res[d]["log_likelihoods_no_dla"], ref["log_likelihoods_no_dla"], rtol=tol, atol=0)
        assert np.array_equal(res[d]["map_inds"], ref["map_inds"])
        assert np.max(np.abs(res[d]["p_dlas"] - ref["p_dlas"])) < P_ATOL
    a6, af = res[6]["sample_log_likelihoods_dla"], res[-1]["sample_log_likelihoods_dla"]
    assert np.max(np.abs(a6 - af) / np.abs(af)) < 2e-12        # 47-bit digits vs FP64 accumulation
    # ragged / masked / short spectra through both paths give the same NaN pattern and values
    sp2 = syn.make_spectra(si["model"], 5, seed=9, dla_fraction=0.5)
    sp2["all_pixel_mask"][1][:] = True
    sp2["all_pixel_mask"][2][::2] = True
    sub = {k: v[::40] for k, v in si["samples"].items()}
    r6 = api.process_qsos(si["model"], sub, sp2, si["prior"], gram_digits=6)
    rf = api.process_qsos(si["model"], sub, sp2, si["prior"], gram_digits=-1)
    assert np.array_equal(np.isnan(r6["sample_log_likelihoods_dla"]), np.isnan(rf["sample_log_likelihoods_dla"]))
    assert np.allclose(r6["sample_log_likelihoods_dla"], rf["sample_log_likelihoods_dla"], rtol=1e-11, atol=0, equal_nan=True)
    assert np.allclose(r6["p_dlas"], rf["p_dlas"], rtol=0, atol=1e-9, equal_nan=True)


def test_state_errors(api, synthetic_inputs):
    from gp_dla_detection_b200._lib import GpdlaError
    si = synthetic_inputs
    bad = dict(si["model"]); bad["M"] = bad["M"][:, :7]
    with pytest.raises(GpdlaError):
        api.DLAProcessor(bad, si["samples"], si["prior"])
    empty = dict(all_wavelengths=[], all_flux=[], all_noise_variance=[], all_pixel_mask=[], z_qsos=np.zeros(0))
    res = api.process_qsos(si["model"], si["samples"], empty, si["prior"])
    assert res["p_dlas"].shape == (0,)


def test_rest_table_agrees_with_direct_evaluation(api, synthetic_inputs):
    """The optical depth from the rest-frame table (default) against the direct evaluation of the line sum
    (rest_table = -1, the arithmetic of voigt.c:282-290 term by term), on both Gram paths and at full sample count:
    the table's polynomials are within 5e-13 of tau / N, which moves the log-likelihoods by < 1e-11 relative."""
    from gp_dla_detection_b200 import synthetic as syn
    from oracle import process_qsos_oracle as O
    si = synthetic_inputs
    sp = syn.make_spectra(si["model"], 4, seed=41, dla_fraction=0.75)
    sp["all_pixel_mask"][2][::3] = True
    ref = O.process_qsos(si["model"], si["samples"], sp, si["prior"], engine="c")
    b = ref["sample_log_likelihoods_dla"]
    for digits in (6, -1):
        tab = api.process_qsos(si["model"], si["samples"], sp, si["prior"], gram_digits=digits, rest_table=0)
        direct = api.process_qsos(si["model"], si["samples"], sp, si["prior"], gram_digits=digits, rest_table=-1)
        a, d = tab["sample_log_likelihoods_dla"], direct["sample_log_likelihoods_dla"]
        assert np.max(np.abs(a - d) / np.abs(d)) < 1e-11, (digits, np.max(np.abs(a - d) / np.abs(d)))
        for r in (tab, direct):
            assert np.max(np.abs(r["sample_log_likelihoods_dla"] - b) / np.abs(b)) < 1e-11
            assert np.array_equal(r["map_inds"], ref["map_inds"])
        assert np.array_equal(tab["log_likelihoods_no_dla"], direct["log_likelihoods_no_dla"])   # a == 1: no table involved
    # a wavelength grid that is NOT the 1e-4-dex BOSS grid (2.5x coarser, irregular): cells are skipped, the table still holds
    rng = np.random.default_rng(8)
    sp2 = syn.make_spectra(si["model"], 2, seed=43, dla_fraction=1.0)
    for q in range(2):
        keep = np.sort(rng.choice(len(sp2["all_flux"][q]), size=len(sp2["all_flux"][q]) * 2 // 5, replace=False))
        for k in ("all_wavelengths", "all_flux", "all_noise_variance", "all_pixel_mask"):
            sp2[k][q] = sp2[k][q][keep]
        sp2["all_wavelengths"][q] = sp2["all_wavelengths"][q] * (1 + 3e-5 * rng.standard_normal(keep.size))
        sp2["all_wavelengths"][q].sort()
    sub = {k: v[::10] for k, v in si["samples"].items()}
    tab = api.process_qsos(si["model"], sub, sp2, si["prior"])
    ref2 = O.process_qsos(si["model"], sub, sp2, si["prior"], engine="c")
    assert_parity(tab, ref2)


def test_zero_noise_variance_pixels(api, synthetic_inputs):
    """A used pixel with noise_variance == 0 (process_qsos.m:194-198 then has d = a^2 omega^2 only): the INT8 path
    has no fixed-point bound for its projection weights there, flags the quasar and leaves it to the FP64 kernels;
    the other quasars of the batch stay on the INT8 path.  Results against the oracle either way."""
    from gp_dla_detection_b200 import synthetic as syn
    from oracle import process_qsos_oracle as O
    si = synthetic_inputs
    sp = syn.make_spectra(si["model"], 4, seed=51, dla_fraction=0.5)
    for q in (1, 3):
        # pixels redward of every sampled DLA (max_z_dla stops 3000 km/s = 43 pixels short of the quasar's Lyman alpha),
        # where the absorption never reaches zero: d = a^2 omega^2 stays positive, as it does for the reference
        nv = sp["all_noise_variance"][q]
        w = np.asarray(sp["all_wavelengths"][q]) / (1 + sp["z_qsos"][q])
        used = np.flatnonzero(~np.asarray(sp["all_pixel_mask"][q]) & (w >= 911.75) & (w <= 1215.75))
        nv[used[[-3, -11, -12, -24]]] = 0.0
    sub = {k: v[::8] for k, v in si["samples"].items()}
    res = api.process_qsos(si["model"], sub, sp, si["prior"])
    ref = O.process_qsos(si["model"], sub, sp, si["prior"], engine="c")
    assert np.all(np.isfinite(res["sample_log_likelihoods_dla"]))
    assert_parity(res, ref)
    # the quasars without such pixels did not change path: bit-identical to a run without the flagged ones
    clean = {k: ([v[0], v[2]] if isinstance(v, list) else np.asarray(v)[[0, 2]]) for k, v in sp.items()}
    res2 = api.process_qsos(si["model"], sub, clean, si["prior"])
    assert np.array_equal(res["sample_log_likelihoods_dla"][[0, 2]], res2["sample_log_likelihoods_dla"])
    # more flagged quasars than the fallback launch has slots
    sp3 = syn.make_spectra(si["model"], 11, seed=52, dla_fraction=0.5)
    for q in range(11):
        w = np.asarray(sp3["all_wavelengths"][q]) / (1 + sp3["z_qsos"][q])
        used = np.flatnonzero(~np.asarray(sp3["all_pixel_mask"][q]) & (w >= 911.75) & (w <= 1215.75))
        sp3["all_noise_variance"][q][used[-2 - q]] = 0.0
    sub3 = {k: v[::50] for k, v in si["samples"].items()}
    assert_parity(api.process_qsos(si["model"], sub3, sp3, si["prior"]), O.process_qsos(si["model"], sub3, sp3, si["prior"], engine="c"))


def test_contexts_of_different_rank_share_a_device(api, synthetic_inputs):
    """Two live contexts with different k (and a third on the other Gram path) interleaved on one GPU give what each
    gives alone: the per-rank tables in constant memory belong to the rank, not to the last context created."""
    from gp_dla_detection_b200 import synthetic as syn
    si = synthetic_inputs
    sub = {k: v[::25] for k, v in si["samples"].items()}
    models = {k: syn.make_model(k) for k in (10, 40, 20)}
    spectra = {k: syn.make_spectra(models[k], 2, seed=60 + k, dla_fraction=0.5) for k in models}
    alone = {k: api.process_qsos(models[k], sub, spectra[k], si["prior"]) for k in models}
    procs = {k: api.DLAProcessor(models[k], sub, si["prior"]) for k in (10, 40, 20)}
    procs["20f"] = api.DLAProcessor(models[20], sub, si["prior"], gram_digits=-1)
    try:
        for _ in range(2):
            for k in (10, 40, 20, 10, 20, 40):
                r = procs[k].process(spectra[k])
                for name in alone[k]:
                    assert np.array_equal(r[name], alone[k][name], equal_nan=True), (k, name)
        rf = procs["20f"].process(spectra[20])
        assert np.allclose(rf["sample_log_likelihoods_dla"], alone[20]["sample_log_likelihoods_dla"], rtol=1e-11, atol=0)
    finally:
        for p in procs.values():
            p.close()


def test_pinned_results_and_full_output_pipeline(api, synthetic_inputs):
    """The host entry returns sample_log_likelihoods_dla batch by batch on its copy stream while the next batch
    computes; with three batches (double-buffered device block reused) and a page-locked destination the rows are
    those of the one-batch run."""
    from gp_dla_detection_b200 import synthetic as syn
    si = synthetic_inputs
    sp = syn.make_spectra(si["model"], 7, seed=71, dla_fraction=0.5)
    sub = {k: v[::20] for k, v in si["samples"].items()}
    one = api.process_qsos(si["model"], sub, sp, si["prior"])
    proc = api.DLAProcessor(si["model"], sub, si["prior"], batch_quasars=3)
    try:
        r = proc.process(sp, pinned_results=True)
        for name in one:
            assert np.array_equal(np.asarray(r[name]), one[name], equal_nan=True), name
        r2 = proc.process(sp, return_sample_log_likelihoods=False)
        assert "sample_log_likelihoods_dla" not in r2 and np.array_equal(r2["p_dlas"], one["p_dlas"], equal_nan=True)
    finally:
        proc.close()


def test_large_batches_use_both_gram_paths(api, synthetic_inputs, monkeypatch):
    """A batch of >= 148 quasars is split: the persistent INT8 kernel's 4-CTA clusters occupy 132 of the 148 SMs, and
    the last ~7 % of the batch run through the FP64 DMMA kernels on the other 16 at the same time (second stream).
    Every quasar must come out exactly as the path that processed it computes it alone, and both within tolerance of
    each other."""
    from gp_dla_detection_b200 import synthetic as syn
    si = synthetic_inputs
    Q = 160
    sp = syn.make_spectra(si["model"], Q, seed=81, dla_fraction=0.3)
    sub = {k: v[::125] for k, v in si["samples"].items()}          # 80 samples: one tile
    split = api.process_qsos(si["model"], sub, sp, si["prior"])
    monkeypatch.setenv("GPDLA_F64_SHARE", "0")
    whole = api.process_qsos(si["model"], sub, sp, si["prior"])
    monkeypatch.delenv("GPDLA_F64_SHARE")
    f64 = api.process_qsos(si["model"], sub, sp, si["prior"], gram_digits=-1)
    n2 = int(round(0.072 * Q))
    for k in split:
        a, w, f = np.asarray(split[k]), np.asarray(whole[k]), np.asarray(f64[k])
        assert np.array_equal(a[:Q - n2], w[:Q - n2], equal_nan=True), k
        assert np.array_equal(a[Q - n2:], f[Q - n2:], equal_nan=True), k
    assert not np.array_equal(split["sample_log_likelihoods_dla"][Q - n2:], whole["sample_log_likelihoods_dla"][Q - n2:])
    assert np.allclose(split["sample_log_likelihoods_dla"], whole["sample_log_likelihoods_dla"], rtol=1e-11, atol=0, equal_nan=True)
    assert np.array_equal(split["map_inds"], whole["map_inds"])
