"""Spectrum preprocessing (read_spec.m:28-38 + preload_qsos.m:18-71): oracle known-answer tests on CPU, parity of
the CUDA kernel against the oracle on the GPU."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def api():
    from gp_dla_detection_b200 import api as A
    return A


def make_raw(Q, seed, z=None):
    """Synthetic SDSS-like coadds: BOSS log-wavelength grid 3.5563 + 1e-4 j, ~4600 pixels, random ivar / masks."""
    rng = np.random.default_rng(seed)
    z_qsos = 2.15 + rng.gamma(2.0, 0.3, Q) if z is None else np.asarray(z, dtype=np.float64)
    raw = dict(flux=[], loglam=[], ivar=[], and_mask=[])
    for q in range(Q):
        n = int(rng.integers(4400, 4650))
        loglam = 3.5563 + 1e-4 * (np.arange(n) + int(rng.integers(0, 30)))
        flux = 3.0 + rng.standard_normal(n)
        ivar = rng.uniform(0.5, 4.0, n)
        ivar[rng.random(n) < 0.03] = 0.0
        am = np.zeros(n, dtype=np.int64)
        am[rng.random(n) < 0.02] |= 1 << 23           # BRIGHTSKY
        am[rng.random(n) < 0.05] |= 1 << 3            # an unrelated bit
        flux[rng.random(n) < 0.002] = np.nan
        for k, v in zip(("flux", "loglam", "ivar", "and_mask"), (flux, loglam, ivar, am)):
            raw[k].append(v)
    return raw, z_qsos


def test_oracle_known_answers():
    from oracle import preload_qsos_oracle as P
    raw, z = make_raw(1, 0, z=[3.2])
    w, f, nv, m = P.read_spec(raw["flux"][0], raw["loglam"][0], raw["ivar"][0], raw["and_mask"][0])
    assert np.array_equal(m, (raw["ivar"][0] == 0) | ((raw["and_mask"][0] & (1 << 23)) != 0))
    assert np.all(np.isinf(nv[raw["ivar"][0] == 0]))
    out = P.preload_qsos(raw, z)
    rest = out["all_wavelengths"][0] / 4.2
    inside = (rest >= 910) & (rest <= 1217)
    # exactly one extra pixel on either side, both unmasked, everything between is kept contiguous
    assert np.count_nonzero(~inside) == 2 and not inside[0] and not inside[-1]
    assert not out["all_pixel_mask"][0][0] and not out["all_pixel_mask"][0][-1]
    # normaliser = median of the unmasked, non-NaN flux in the rest-frame window [1310, 1325]
    r = w / 4.2
    sel = f[(r >= 1310) & (r <= 1325) & ~m]; sel = sel[~np.isnan(sel)]
    assert out["all_normalizers"][0] == np.median(sel)
    keep = np.isin(w, out["all_wavelengths"][0])
    assert np.array_equal(out["all_flux"][0], f[keep] / np.median(sel), equal_nan=True)
    # flags: everything masked in the normalisation window -> bit 3; too few pixels -> bit 4; input flag -> skipped
    raw2, z2 = make_raw(3, 1, z=[2.5, 2.5, 2.5])
    r2 = 10 ** raw2["loglam"][0] / 3.5
    raw2["ivar"][0][(r2 >= 1310) & (r2 <= 1325)] = 0.0
    r3 = 10 ** raw2["loglam"][1] / 3.5
    idx = np.flatnonzero((r3 >= 911.75) & (r3 <= 1215.75))
    raw2["ivar"][1][idx[150:]] = 0.0
    out2 = P.preload_qsos(raw2, z2, filter_flags=[0, 0, 1])
    assert list(out2["filter_flags"]) == [4, 8, 1]
    assert all(len(out2["all_flux"][q]) == 0 for q in range(3)) and np.all(out2["all_normalizers"] == 0)


@pytest.mark.gpu
def test_preload_matches_oracle(api):
    from oracle import preload_qsos_oracle as P
    raw, z = make_raw(24, 7)
    # edge cases: unnormalisable, too few pixels, pre-flagged, masked pixels right outside the loading window
    r0 = 10 ** raw["loglam"][0] / (1 + z[0]); raw["ivar"][0][(r0 >= 1310) & (r0 <= 1325)] = 0.0
    r1 = 10 ** raw["loglam"][1] / (1 + z[1]); i1 = np.flatnonzero((r1 >= 911.75) & (r1 <= 1215.75)); raw["ivar"][1][i1[100:]] = 0.0
    flags = np.zeros(24, dtype=np.uint8); flags[2] = 2
    r3 = 10 ** raw["loglam"][3] / (1 + z[3]); i3 = np.flatnonzero((r3 >= 910) & (r3 <= 1217))
    raw["ivar"][3][i3[-1] + 1:i3[-1] + 4] = 0.0; raw["ivar"][3][i3[0] - 3:i3[0]] = 0.0
    ref = P.preload_qsos(raw, z, filter_flags=flags.copy())
    res = api.preload_qsos(raw, z, filter_flags=flags)
    assert np.array_equal(res["filter_flags"], ref["filter_flags"])
    assert list(res["filter_flags"][:3]) == [4, 8, 2]
    assert np.array_equal(res["all_normalizers"], ref["all_normalizers"])        # order statistics: bit-exact
    for q in range(24):
        assert len(res["all_flux"][q]) == len(ref["all_flux"][q]), q
        assert np.array_equal(res["all_pixel_mask"][q], ref["all_pixel_mask"][q])
        assert np.allclose(res["all_wavelengths"][q], ref["all_wavelengths"][q], rtol=4e-16, atol=0)   # exp10 vs 10**x: 2 ulp
        assert np.array_equal(res["all_flux"][q], ref["all_flux"][q], equal_nan=True)                  # one IEEE division
        assert np.array_equal(res["all_noise_variance"][q], ref["all_noise_variance"][q])
    # the edge pixel after the window skips the three masked pixels that follow it
    assert np.isclose(res["all_wavelengths"][3][-1], 10 ** raw["loglam"][3][i3[-1] + 4], rtol=1e-15)
    assert len(res["all_flux"][3]) == len(i3) + (2 if i3[0] > 3 else 1)
    # L_out too small is an error, not a truncation
    from gp_dla_detection_b200._lib import GpdlaError
    with pytest.raises(GpdlaError):
        api.preload_qsos(raw, z, L_out=500)


@pytest.mark.gpu
def test_preload_device_feeds_process_device(api, synthetic_inputs):
    """Raw coadds -> device preprocessing -> process_device, without a host round trip, equals host preprocessing
    (oracle) -> process_qsos."""
    import torch
    from oracle import preload_qsos_oracle as P
    si = synthetic_inputs
    raw, z = make_raw(4, 11)
    for q in range(4):                       # NaN flux only in masked pixels, as in real coadds
        raw["ivar"][q][np.isnan(raw["flux"][q])] = 0.0
    pad = api._pad_raw(raw)
    t = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in pad.items()}
    zt = torch.from_numpy(z).cuda()
    pre = api.preload_qsos_device(t["flux"], t["loglam"], t["ivar"], t["and_mask"], t["lengths"], zt)
    sub = {k: v[::100] for k, v in si["samples"].items()}
    proc = api.DLAProcessor(si["model"], sub, si["prior"])
    dev = proc.process_device(pre["wavelengths"], pre["flux"], pre["noise_variance"], pre["pixel_mask"], pre["lengths"], zt)
    torch.cuda.synchronize()
    ref_sp = P.preload_qsos(raw, z)
    ref_sp["z_qsos"] = z
    ref = api.process_qsos(si["model"], sub, ref_sp, si["prior"])
    assert np.all(np.isfinite(ref["log_likelihoods_no_dla"]))
    assert np.allclose(dev["log_likelihoods_no_dla"].cpu().numpy(), ref["log_likelihoods_no_dla"], rtol=1e-9)
    assert np.allclose(dev["p_dlas"].cpu().numpy(), ref["p_dlas"], atol=1e-8)
