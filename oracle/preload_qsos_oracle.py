"""TEST INFRASTRUCTURE (oracle): literal NumPy restatement of the reference's spectrum preprocessing,
read_spec.m:28-38 + preload_qsos.m:18-71, one quasar at a time in the reference's operation order.
Parity unpinned against reference outputs (no MATLAB/Octave/FITS data in the image): pinned by the known-answer
tests in tests/test_preload.py only."""
import numpy as np

# set_parameters.m:21-34
loading_min_lambda, loading_max_lambda = 910.0, 1217.0
normalization_min_lambda, normalization_max_lambda = 1310.0, 1325.0
min_lambda, max_lambda = 911.75, 1215.75
min_num_pixels = 200
BRIGHTSKY = 24   # read_spec.m:9


def read_spec(flux, loglam, ivar, and_mask):
    """read_spec.m:28-38 on the four table columns."""
    wavelengths = 10.0 ** np.asarray(loglam, dtype=np.float64)                       # :28
    with np.errstate(divide="ignore"):
        noise_variance = 1.0 / np.asarray(ivar, dtype=np.float64)                    # :31
    pixel_mask = (np.asarray(ivar) == 0) | (((np.asarray(and_mask).astype(np.int64) >> (BRIGHTSKY - 1)) & 1) == 1)   # :35-37
    return wavelengths, np.asarray(flux, dtype=np.float64), noise_variance, pixel_mask


def preload_qsos(raw, z_qsos, filter_flags=None):
    """raw: dict of ragged lists flux / loglam / ivar / and_mask.  Returns the variables saved at preload_qsos.m:73-79
    (empty arrays for skipped quasars) and the updated filter flags."""
    Q = len(z_qsos)
    flags = np.zeros(Q, dtype=np.uint8) if filter_flags is None else np.array(filter_flags, dtype=np.uint8)
    out = dict(all_wavelengths=[np.zeros(0)] * Q, all_flux=[np.zeros(0)] * Q, all_noise_variance=[np.zeros(0)] * Q,
               all_pixel_mask=[np.zeros(0, dtype=bool)] * Q, all_normalizers=np.zeros(Q), filter_flags=flags)
    for i in range(Q):
        if flags[i] > 0:                                                              # :19-21
            continue
        w, f, nv, m = read_spec(raw["flux"][i], raw["loglam"][i], raw["ivar"][i], raw["and_mask"][i])
        rest = w / (1.0 + z_qsos[i])                                                  # :26, set_parameters.m:14-15
        ind = (rest >= normalization_min_lambda) & (rest <= normalization_max_lambda) & ~m   # :29-31
        sel = f[ind]; sel = sel[~np.isnan(sel)]
        if sel.size == 0:                                                             # :33-39 nanmedian = NaN
            flags[i] |= 4
            continue
        this_median = np.median(sel)
        ind = (rest >= min_lambda) & (rest <= max_lambda) & ~m                        # :41-43
        if np.count_nonzero(ind) < min_num_pixels:                                    # :46-49
            flags[i] |= 8
            continue
        out["all_normalizers"][i] = this_median
        f = f / this_median                                                           # :51
        nv = nv / this_median ** 2                                                    # :52
        ind = (rest >= loading_min_lambda) & (rest <= loading_max_lambda)             # :54-55
        available = np.flatnonzero(~ind & ~m)                                         # :58
        inside = np.flatnonzero(ind)
        after = available[available > inside[-1]]                                     # :59
        before = available[available < inside[0]]                                     # :60
        if after.size:
            ind[after.min()] = True
        if before.size:
            ind[before.max()] = True
        out["all_wavelengths"][i], out["all_flux"][i] = w[ind], f[ind]                # :64-67
        out["all_noise_variance"][i], out["all_pixel_mask"][i] = nv[ind], m[ind]
    return out
