"""Host-side mirror of the reference's interface for the hot path.

``process_qsos`` takes what the MATLAB script ``process_qsos.m`` reads from its ``.mat`` inputs
(process_qsos.m:4-49) and returns the variables it saves (process_qsos.m:236-244) with the
same names and shapes, plus ``map_z_dlas`` / ``map_log_nhis`` / ``map_inds``
(generate_ascii_catalog.m:73-80).  ``voigt`` mirrors the MEX signature
``profile = voigt(lambdas, z, N [, num_lines])`` (voigt.c:8-13,253-304; default 31 lines).

All compute happens in libgpdla.so (hand-written sm_100a CUDA) through the C ABI of
``include/gpdla.h``.  PyTorch is used only for device buffers and streams in the
device-resident entry points.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Sequence

import numpy as np

from . import _lib
from .params import DEFAULT, Parameters

RESULT_NAMES = ["min_z_dlas", "max_z_dlas", "log_priors_no_dla", "log_priors_dla", "log_likelihoods_no_dla",
                "log_likelihoods_dla", "log_posteriors_no_dla", "log_posteriors_dla", "model_posteriors",
                "p_no_dlas", "p_dlas", "map_z_dlas", "map_log_nhis"]


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _dp(a: np.ndarray):
    return a.ctypes.data_as(_lib.c_double_p)


def pad_spectra(spectra: Dict) -> Dict[str, np.ndarray]:
    """Ragged cell arrays (preload_qsos.m:73-79) -> padded ``[Q x L_max]`` planes + lengths."""
    if "lengths" in spectra:
        return spectra
    W = spectra["all_wavelengths"]
    Q = len(W)
    lengths = np.array([len(w) for w in W], dtype=np.int32)
    L_max = max(int(lengths.max()) if Q else 1, 1)
    out = dict(wavelengths=np.zeros((Q, L_max)), flux=np.zeros((Q, L_max)),
               noise_variance=np.ones((Q, L_max)), pixel_mask=np.ones((Q, L_max), dtype=np.uint8),
               lengths=lengths, z_qsos=_f64(spectra["z_qsos"]))
    for q in range(Q):
        n = lengths[q]
        out["wavelengths"][q, :n] = spectra["all_wavelengths"][q]
        out["flux"][q, :n] = spectra["all_flux"][q]
        out["noise_variance"][q, :n] = spectra["all_noise_variance"][q]
        out["pixel_mask"][q, :n] = np.asarray(spectra["all_pixel_mask"][q]).astype(np.uint8)
    return out


def _check_planes(W, F, V, Mk, lengths, z):
    """The padded planes must share one [Q x L_max] shape and every length must fit its row."""
    if W.ndim != 2 or F.shape != W.shape or V.shape != W.shape or Mk.shape != W.shape:
        raise ValueError("wavelengths, flux, noise_variance and pixel_mask must share one (Q, L_max) shape")
    if lengths.shape != (W.shape[0],) or z.shape != (W.shape[0],):
        raise ValueError("lengths and z_qsos must have one entry per quasar")
    if lengths.size and (lengths.min() < 0 or lengths.max() > W.shape[1]):
        raise ValueError("lengths must lie in [0, L_max]")


class DLAProcessor:
    """A libgpdla context holding the learned null model, the DLA samples and the prior
    catalogue on one GPU (what process_qsos.m:4-40 loads once before its quasar loop)."""

    def __init__(self, model: Dict, samples: Dict, prior: Dict, params: Parameters = DEFAULT, device: int = 0,
                 batch_quasars: int = 0, gram_digits: int = 0, rest_table: int = 0):
        """``gram_digits``: arithmetic of the Gram contraction -- 0 default (exact-product INT8 tensor-core path
        with 6 digits for k = 20 and 40, FP64 DMMA for k = 10), -1 FP64 DMMA, 5 / 6 INT8 path with that many digits (k = 40: 6).
        ``rest_table``: 0 default (optical depth from the rest-frame table away from the line centres), -1 direct
        evaluation of the line sum everywhere (``gpdla_params.rest_table``)."""
        self._lib = _lib.load()
        self._pinned = {}
        self._ctx = ctypes.c_void_p()
        _lib.check(self._lib.gpdla_create(ctypes.byref(self._ctx), int(device)))
        self.device = int(device)
        self.params = params
        p = _lib.GpdlaParams()
        self._lib.gpdla_default_parameters(ctypes.byref(p))
        p.min_lambda, p.max_lambda = params.min_lambda, params.max_lambda
        p.prior_z_qso_increase = params.prior_z_qso_increase
        p.min_z_cut, p.max_z_cut = params.min_z_cut, params.max_z_cut
        p.pixel_spacing, p.num_lines, p.batch_quasars = params.pixel_spacing, params.num_lines, int(batch_quasars)
        p.gram_digits = int(gram_digits)
        p.rest_table = int(rest_table)
        _lib.check(self._lib.gpdla_set_parameters(self._ctx, ctypes.byref(p)), self._ctx)
        rest, mu, M, lw = (_f64(model[k]) for k in ("rest_wavelengths", "mu", "M", "log_omega"))
        if M.shape != (rest.size, M.shape[1]) or mu.size != rest.size or lw.size != rest.size:
            raise ValueError("model arrays must be rest_wavelengths (n,), mu (n,), M (n, k), log_omega (n,)")
        self.k = int(M.shape[1])
        _lib.check(self._lib.gpdla_set_model(self._ctx, _dp(rest), rest.size, _dp(mu), _dp(M), self.k, _dp(lw),
                                             float(model["log_c_0"]), float(model["log_tau_0"]),
                                             float(model["log_beta"])), self._ctx)
        off, lnhi, nhi = (_f64(samples[k]).ravel() for k in ("offset_samples", "log_nhi_samples", "nhi_samples"))
        self.num_dla_samples = int(off.size)
        _lib.check(self._lib.gpdla_set_samples(self._ctx, _dp(off), _dp(lnhi), _dp(nhi), off.size), self._ctx)
        self.has_lls = "lls_nhi_samples" in samples and "Z_lls" in samples and "Z_dla" in samples
        if self.has_lls:   # multi_dlas/set_lls_parameters.m:55-71
            lls = _f64(samples["lls_nhi_samples"]).ravel()
            _lib.check(self._lib.gpdla_set_lls_samples(self._ctx, _dp(lls), lls.size, float(samples["Z_lls"]),
                                                       float(samples["Z_dla"])), self._ctx)
        pz = _f64(prior["z_qsos"]).ravel()
        pd = np.ascontiguousarray(np.asarray(prior["dla_ind"]).astype(np.uint8)).ravel()
        _lib.check(self._lib.gpdla_set_prior(self._ctx, _dp(pz), pd.ctypes.data_as(_lib.c_u8_p), pz.size), self._ctx)

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx:
            self._lib.gpdla_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()
        for ptr, _ in getattr(self, "_pinned", {}).values():
            self._lib.gpdla_host_free(ptr)
        self._pinned = {}

    def _pinned_array(self, name: str, shape, dtype=np.float64) -> np.ndarray:
        """A page-locked host array owned by this processor (``gpdla_host_alloc``), reused while the shape
        holds: result copies into it overlap the next batch's kernels.  Valid until the next call / ``close``."""
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        have = self._pinned.get(name)
        if have is None or have[1] < nbytes:
            if have is not None:
                self._lib.gpdla_host_free(have[0])
            ptr = ctypes.c_void_p()
            _lib.check(self._lib.gpdla_host_alloc(ctypes.byref(ptr), max(nbytes, 1)))
            self._pinned[name] = have = (ptr, nbytes)
        buf = (ctypes.c_char * max(nbytes, 1)).from_address(have[0].value)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launch_count(self) -> int:
        return int(self._lib.gpdla_launch_count(self._ctx))

    def set_profiling(self, enable: bool):
        _lib.check(self._lib.gpdla_set_profiling(self._ctx, int(bool(enable))), self._ctx)

    def profile_read(self):
        """(summed ms, launches) of the fused log-likelihood kernel since the last read."""
        ms, n = ctypes.c_double(), ctypes.c_int64()
        _lib.check(self._lib.gpdla_profile_read(self._ctx, ctypes.byref(ms), ctypes.byref(n)), self._ctx)
        return ms.value, n.value

    # ---------------------------------------------------------------- host buffers (the drop-in call)
    def process(self, spectra: Dict, return_sample_log_likelihoods: bool = True,
                pinned_results: bool = False) -> Dict[str, np.ndarray]:
        """The body of process_qsos.m for the given spectra (host arrays in, host arrays out).
        ``pinned_results``: ``sample_log_likelihoods_dla`` (80 KB per quasar) is returned in a page-locked buffer
        owned by this processor -- valid until the next ``process`` call -- so that its copy overlaps compute."""
        sp = pad_spectra(spectra)
        W, F, V = _f64(sp["wavelengths"]), _f64(sp["flux"]), _f64(sp["noise_variance"])
        Mk = np.ascontiguousarray(sp["pixel_mask"], dtype=np.uint8)
        lengths = np.ascontiguousarray(sp["lengths"], dtype=np.int32).ravel()
        z = _f64(sp["z_qsos"]).ravel()
        _check_planes(W, F, V, Mk, lengths, z)
        Q, L_max = W.shape
        S = self.num_dla_samples
        out = {n: np.full((Q, 2) if n == "model_posteriors" else (Q,), np.nan) for n in RESULT_NAMES}
        out["map_inds"] = np.full(Q, -1, dtype=np.int64)
        if return_sample_log_likelihoods:
            if pinned_results:
                out["sample_log_likelihoods_dla"] = self._pinned_array("sll", (Q, S))
                out["sample_log_likelihoods_dla"].fill(np.nan)
            else:
                out["sample_log_likelihoods_dla"] = np.full((Q, S), np.nan)
        res = _lib.GpdlaResults()
        for n in _lib.RESULT_F64:
            setattr(res, n, out[n].ctypes.data if n in out else None)
        res.map_inds = out["map_inds"].ctypes.data
        vp = lambda a: ctypes.c_void_p(a.ctypes.data)
        _lib.check(self._lib.gpdla_process_qsos(self._ctx, Q, L_max, vp(W), vp(F), vp(V), vp(Mk), vp(lengths),
                                                vp(z), ctypes.byref(res)), self._ctx)
        return out

    # ---------------------------------------------------------------- multi-DLA + sub-DLA + mean flux
    def process_multi(self, spectra: Dict, max_dlas: int = 4, base_sample_inds: Optional[np.ndarray] = None,
                      return_samples: bool = True) -> Dict[str, np.ndarray]:
        """``process_qsos_multiple_dlas_meanflux`` (multi_dlas/process_qsos_multiple_dlas_meanflux.m): results
        with the reference's names and shapes -- ``sample_log_likelihoods_dla`` (Q, S, max_dlas),
        ``base_sample_inds`` (Q, S, max_dlas-1; 0-based), ``MAP_*`` (Q, max_dlas, max_dlas),
        ``model_posteriors`` (Q, 2 + max_dlas).  ``base_sample_inds`` given as input (same shape) replaces
        the built-in resampling (rng('default') + randsample) for parity runs."""
        if not self.has_lls:
            raise ValueError("samples must hold lls_nhi_samples, Z_lls and Z_dla for the multi-DLA path")
        sp = pad_spectra(spectra)
        W, F, V = _f64(sp["wavelengths"]), _f64(sp["flux"]), _f64(sp["noise_variance"])
        Mk = np.ascontiguousarray(sp["pixel_mask"], dtype=np.uint8)
        lengths = np.ascontiguousarray(sp["lengths"], dtype=np.int32).ravel()
        z = _f64(sp["z_qsos"]).ravel()
        _check_planes(W, F, V, Mk, lengths, z)
        Q, L_max = W.shape
        S, MD = self.num_dla_samples, int(max_dlas)
        shapes = dict(log_priors_dla=(Q, MD), log_likelihoods_dla=(Q, MD), log_posteriors_dla=(Q, MD),
                      model_posteriors=(Q, MD + 2), MAP_z_dlas=(Q, MD, MD), MAP_log_nhis=(Q, MD, MD))
        out = {}
        for n in _lib.MULTI_RESULT_FIELDS[:17]:
            out[n] = np.full(shapes.get(n, (Q,)), np.nan)
        out["MAP_inds"] = np.full((Q, MD, MD), -1, dtype=np.int64)
        if return_samples:
            out["sample_log_likelihoods_dla"] = np.full((Q, MD, S), np.nan)
            out["sample_log_likelihoods_lls"] = np.full((Q, S), np.nan)
            out["base_sample_inds"] = np.zeros((Q, max(MD - 1, 0), S), dtype=np.int32)
        res = _lib.GpdlaMultiResults()
        for n in _lib.MULTI_RESULT_FIELDS:
            setattr(res, n, out[n].ctypes.data if n in out and out[n].size else None)
        bin_ = None
        if base_sample_inds is not None:
            bin_ = np.ascontiguousarray(np.transpose(np.asarray(base_sample_inds), (0, 2, 1)), dtype=np.int32)
            assert bin_.shape == (Q, MD - 1, S)
        vp = lambda a: ctypes.c_void_p(a.ctypes.data)
        _lib.check(self._lib.gpdla_process_qsos_multi(
            self._ctx, Q, L_max, vp(W), vp(F), vp(V), vp(Mk), vp(lengths), vp(z), MD,
            vp(bin_) if bin_ is not None else None, ctypes.byref(res)), self._ctx)
        if return_samples:   # reference layout: (quasar, sample, level)
            out["sample_log_likelihoods_dla"] = np.transpose(out["sample_log_likelihoods_dla"], (0, 2, 1))
            out["base_sample_inds"] = np.transpose(out["base_sample_inds"], (0, 2, 1))
        return out

    # ---------------------------------------------------------------- device-resident buffers
    def process_device(self, wavelengths, flux, noise_variance, pixel_mask, lengths, z_qsos,
                       return_sample_log_likelihoods: bool = False, out: Optional[Dict] = None) -> Dict:
        """Inputs are CUDA tensors (float64 [Q x L_max], uint8 mask, int32 lengths, float64 z);
        returns CUDA tensors.  Asynchronous on torch's current stream."""
        import torch
        Q, L_max = wavelengths.shape
        dev = wavelengths.device
        assert dev.type == "cuda" and dev.index == self.device
        for t, dt in ((wavelengths, torch.float64), (flux, torch.float64), (noise_variance, torch.float64),
                      (pixel_mask, torch.uint8), (lengths, torch.int32), (z_qsos, torch.float64)):
            assert t.is_cuda and t.dtype == dt and t.is_contiguous()
        if out is None:
            out = {n: torch.empty((Q, 2) if n == "model_posteriors" else (Q,), dtype=torch.float64, device=dev)
                   for n in RESULT_NAMES}
            out["map_inds"] = torch.empty(Q, dtype=torch.int64, device=dev)
            if return_sample_log_likelihoods:
                out["sample_log_likelihoods_dla"] = torch.empty((Q, self.num_dla_samples), dtype=torch.float64,
                                                                device=dev)
        res = _lib.GpdlaResults()
        for n in _lib.RESULT_F64:
            setattr(res, n, out[n].data_ptr() if n in out else None)
        res.map_inds = out["map_inds"].data_ptr()
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(self._lib.gpdla_process_qsos_device(
            self._ctx, Q, L_max, wavelengths.data_ptr(), flux.data_ptr(), noise_variance.data_ptr(),
            pixel_mask.data_ptr(), lengths.data_ptr(), z_qsos.data_ptr(), ctypes.byref(res),
            ctypes.c_void_p(stream)), self._ctx)
        return out


def process_qsos(model: Dict, samples: Dict, spectra: Dict, prior: Dict, params: Parameters = DEFAULT,
                 device: int = 0, return_sample_log_likelihoods: bool = True, gram_digits: int = 0,
                 rest_table: int = 0) -> Dict[str, np.ndarray]:
    """Run the DLA detection algorithm on the given objects (process_qsos.m)."""
    proc = DLAProcessor(model, samples, prior, params, device, gram_digits=gram_digits, rest_table=rest_table)
    try:
        return proc.process(spectra, return_sample_log_likelihoods)
    finally:
        proc.close()


def process_qsos_multiple_dlas_meanflux(model: Dict, samples: Dict, spectra: Dict, prior: Dict, max_dlas: int = 4,
                                        params: Parameters = DEFAULT, device: int = 0,
                                        base_sample_inds: Optional[np.ndarray] = None,
                                        return_samples: bool = True, batch_quasars: int = 0,
                                        gram_digits: int = 0, rest_table: int = 0) -> Dict[str, np.ndarray]:
    """Multi-DLA / sub-DLA / mean-flux processing (multi_dlas/process_qsos_multiple_dlas_meanflux.m).
    ``samples`` additionally holds ``lls_nhi_samples``, ``Z_lls``, ``Z_dla`` (set_lls_parameters.m)."""
    proc = DLAProcessor(model, samples, prior, params, device, batch_quasars=batch_quasars, gram_digits=gram_digits,
                        rest_table=rest_table)
    try:
        return proc.process_multi(spectra, max_dlas, base_sample_inds, return_samples)
    finally:
        proc.close()


def _pad_raw(raw: Dict) -> Dict[str, np.ndarray]:
    F = raw["flux"]
    Q = len(F)
    lengths = np.array([len(f) for f in F], dtype=np.int32)
    L = max(int(lengths.max()) if Q else 1, 1)
    out = dict(flux=np.zeros((Q, L)), loglam=np.zeros((Q, L)), ivar=np.zeros((Q, L)),
               and_mask=np.zeros((Q, L), dtype=np.int32), lengths=lengths)
    for q in range(Q):
        n = lengths[q]
        out["flux"][q, :n] = raw["flux"][q]
        out["loglam"][q, :n] = raw["loglam"][q]
        out["ivar"][q, :n] = raw["ivar"][q]
        out["and_mask"][q, :n] = np.asarray(raw["and_mask"][q]).astype(np.int64).astype(np.int32)
    return out


def preload_qsos(raw: Dict, z_qsos, filter_flags=None, params: Parameters = DEFAULT, device: int = 0,
                 L_out: int = 0) -> Dict:
    """Spectrum preprocessing on the GPU: read_spec.m:28-38 + preload_qsos.m:18-71.

    ``raw`` holds the four columns of the SDSS speclite coadd table per quasar (ragged lists ``flux``, ``loglam``,
    ``ivar``, ``and_mask``; read_spec.m:11-26).  Returns the variables preload_qsos.m:73-79 saves --
    ``all_wavelengths``, ``all_flux``, ``all_noise_variance``, ``all_pixel_mask`` (ragged lists, empty for skipped
    quasars), ``all_normalizers`` -- and the updated ``filter_flags``; the dict can be passed to ``process_qsos``
    as ``spectra`` once ``z_qsos`` is added and the flagged quasars are dropped."""
    import torch
    lib = _lib.load()
    pad = _pad_raw(raw)
    Q, L_in = pad["flux"].shape
    z = _f64(z_qsos)
    if L_out <= 0:
        # the loading window spans log10(1217 / 910) / 1e-4 = 1263 pixels of the BOSS grid, + 2 edge pixels
        L_out = min(L_in, 1280)
    p = _lib.GpdlaPreloadParams()
    lib.gpdla_default_preload_parameters(ctypes.byref(p))
    p.min_lambda, p.max_lambda = params.min_lambda, params.max_lambda
    p.loading_min_lambda, p.loading_max_lambda = params.loading_min_lambda, params.loading_max_lambda
    p.min_num_pixels = params.min_num_pixels
    flags_in = None if filter_flags is None else np.ascontiguousarray(filter_flags, dtype=np.uint8)
    out = dict(wavelengths=np.zeros((Q, L_out)), flux=np.zeros((Q, L_out)), noise_variance=np.zeros((Q, L_out)),
               pixel_mask=np.zeros((Q, L_out), dtype=np.uint8), lengths=np.zeros(Q, dtype=np.int32),
               normalizers=np.zeros(Q), filter_flags=np.zeros(Q, dtype=np.uint8))
    vp = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
    with torch.cuda.device(device):
        _lib.check(lib.gpdla_preload_qsos(Q, L_in, vp(pad["flux"]), vp(pad["loglam"]), vp(pad["ivar"]), vp(pad["and_mask"]),
                                          vp(pad["lengths"]), vp(z), vp(flags_in), ctypes.byref(p), L_out,
                                          vp(out["wavelengths"]), vp(out["flux"]), vp(out["noise_variance"]),
                                          vp(out["pixel_mask"]), vp(out["lengths"]), vp(out["normalizers"]),
                                          vp(out["filter_flags"])))
    n = out["lengths"]
    return dict(all_wavelengths=[out["wavelengths"][q, :n[q]].copy() for q in range(Q)],
                all_flux=[out["flux"][q, :n[q]].copy() for q in range(Q)],
                all_noise_variance=[out["noise_variance"][q, :n[q]].copy() for q in range(Q)],
                all_pixel_mask=[out["pixel_mask"][q, :n[q]].astype(bool) for q in range(Q)],
                all_normalizers=out["normalizers"], filter_flags=out["filter_flags"])


def preload_qsos_device(flux, loglam, ivar, and_mask, lengths, z_qsos, filter_flags=None, params: Parameters = DEFAULT,
                        L_out: int = 1280, stream: int = 0):
    """Device-resident form: torch CUDA tensors in ([Q x L_in] planes), padded [Q x L_out] planes + lengths out --
    exactly the arguments of ``DLAProcessor.process_device`` -- so real spectra never return to the host."""
    import torch
    lib = _lib.load()
    Q, L_in = flux.shape
    dev = flux.device
    out = dict(wavelengths=torch.empty((Q, L_out), dtype=torch.float64, device=dev),
               flux=torch.empty((Q, L_out), dtype=torch.float64, device=dev),
               noise_variance=torch.empty((Q, L_out), dtype=torch.float64, device=dev),
               pixel_mask=torch.empty((Q, L_out), dtype=torch.uint8, device=dev),
               lengths=torch.empty(Q, dtype=torch.int32, device=dev),
               normalizers=torch.empty(Q, dtype=torch.float64, device=dev),
               filter_flags=torch.empty(Q, dtype=torch.uint8, device=dev))
    p = _lib.GpdlaPreloadParams()
    lib.gpdla_default_preload_parameters(ctypes.byref(p))
    p.min_lambda, p.max_lambda = params.min_lambda, params.max_lambda
    p.loading_min_lambda, p.loading_max_lambda = params.loading_min_lambda, params.loading_max_lambda
    p.min_num_pixels = params.min_num_pixels
    ptr = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
    with torch.cuda.device(dev):
        _lib.check(lib.gpdla_preload_qsos_device(Q, L_in, ptr(flux), ptr(loglam), ptr(ivar), ptr(and_mask), ptr(lengths),
                                                 ptr(z_qsos), ptr(filter_flags), ctypes.byref(p), L_out,
                                                 ptr(out["wavelengths"]), ptr(out["flux"]), ptr(out["noise_variance"]),
                                                 ptr(out["pixel_mask"]), ptr(out["lengths"]), ptr(out["normalizers"]),
                                                 ptr(out["filter_flags"]), ctypes.c_void_p(stream)))
    return out


def _forest_arrays(num_forest_lines, all_transition_wavelengths, all_oscillator_strengths):
    nl = int(num_forest_lines)
    if nl == 0:
        return 0, None, None
    tw, osc = _f64(all_transition_wavelengths).ravel(), _f64(all_oscillator_strengths).ravel()
    if not (1 <= nl <= _lib.MAX_LINES) or tw.size < nl or osc.size < nl:
        raise ValueError("num_forest_lines must be 1..31 with that many transition wavelengths and oscillator strengths")
    return nl, tw, osc


def objective_lyseries(x, centered_rest_fluxes, lya_1pzs, rest_noise_variances, num_forest_lines,
                       all_transition_wavelengths, all_oscillator_strengths, device: int = 0):
    """``[f, g] = objective_lyseries(x, centered_rest_fluxes, lya_1pzs, rest_noise_variances, num_forest_lines,
    all_transition_wavelengths, all_oscillator_strengths)`` of multi_dlas/objective_lyseries.m:13-87 on the GPU: the
    training objective with the Lyman-series effective optical depth (multi_dlas/spectrum_loss_lyseries.m:20-47).
    ``num_forest_lines = 0`` is ``objective``."""
    import torch
    lib = _lib.load()
    y, z1, nv, xx = _f64(centered_rest_fluxes), _f64(lya_1pzs), _f64(rest_noise_variances), _f64(x)
    N, P = y.shape
    k = (xx.size - 3) // P - 1                                            # objective.m:17
    if xx.size != P * (k + 1) + 3 or z1.shape != y.shape or nv.shape != y.shape:
        raise ValueError("x must have num_pixels * (k + 1) + 3 entries and the three data matrices one shape")
    nl, tw, osc = _forest_arrays(num_forest_lines, all_transition_wavelengths, all_oscillator_strengths)
    f = np.zeros(1); g = np.zeros(xx.size)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    with torch.cuda.device(device):
        _lib.check(lib.gpdla_objective_lyseries(N, P, k, vp(y), vp(z1), vp(nv), nl, _dp(tw) if nl else None,
                                                _dp(osc) if nl else None, vp(xx), vp(f), vp(g)))
    return float(f[0]), g


def objective(x, centered_rest_fluxes, lya_1pzs, rest_noise_variances, device: int = 0):
    """``[f, g] = objective(x, centered_rest_fluxes, lya_1pzs, rest_noise_variances)`` of objective.m:12-73 on the
    GPU: negative log-likelihood of the training set and its gradient, ``x = [M(:); log_omega; log c_0; log tau_0;
    log beta]`` (M column-major as in MATLAB).  Host arrays in, ``(float, ndarray)`` out."""
    return objective_lyseries(x, centered_rest_fluxes, lya_1pzs, rest_noise_variances, 0, None, None, device)


class TrainingObjective:
    """The training matrices of learn_qso_model.m:36-75 resident on one GPU; ``ev(x) -> (f, g)`` evaluates
    objective.m for a parameter vector (what minFunc calls at learn_qso_model.m:97-99), uploading only ``x``."""

    def __init__(self, centered_rest_fluxes, lya_1pzs, rest_noise_variances, k: int, device: int = 0,
                 num_forest_lines: int = 0, all_transition_wavelengths=None, all_oscillator_strengths=None):
        """``num_forest_lines > 0`` selects objective_lyseries.m (learn_qso_model_meanflux.m:140-142)."""
        import torch
        self.nl, self.tw, self.osc = _forest_arrays(num_forest_lines, all_transition_wavelengths, all_oscillator_strengths)
        self._lib = _lib.load()
        self._torch = torch
        self.dev = torch.device("cuda", device)
        self.N, self.P = np.shape(centered_rest_fluxes)
        self.k = int(k)
        up = lambda a: torch.from_numpy(_f64(a)).to(self.dev)
        self.y, self.z1, self.nv = up(centered_rest_fluxes), up(lya_1pzs), up(rest_noise_variances)
        self.nx = self.P * (self.k + 1) + 3
        self.x = torch.empty(self.nx, dtype=torch.float64, device=self.dev)
        self.g = torch.empty(self.nx, dtype=torch.float64, device=self.dev)
        self.f = torch.empty(1, dtype=torch.float64, device=self.dev)

    def evaluate_device(self, x_dev, stream: int = 0):
        """x on the device -> (f, g) device tensors (asynchronous)."""
        ptr = lambda t: ctypes.c_void_p(t.data_ptr())
        with self._torch.cuda.device(self.dev):
            _lib.check(self._lib.gpdla_objective_lyseries_device(
                self.N, self.P, self.k, ptr(self.y), ptr(self.z1), ptr(self.nv), self.nl, _dp(self.tw) if self.nl else None,
                _dp(self.osc) if self.nl else None, ptr(x_dev), ptr(self.f), ptr(self.g), ctypes.c_void_p(stream)))
        return self.f, self.g

    def __call__(self, x):
        xx = _f64(x)
        if xx.size != self.nx:
            raise ValueError("x must have %d entries" % self.nx)
        self.x.copy_(self._torch.from_numpy(xx))
        f, g = self.evaluate_device(self.x)
        return float(f.item()), g.cpu().numpy()

    def close(self):
        self.y = self.z1 = self.nv = self.x = self.g = self.f = None


def learn_qso_model(centered_rest_fluxes, lya_1pzs, rest_noise_variances, initial_x, k: int, max_iter: int = 2000,
                    device: int = 0):
    """The optimisation of learn_qso_model.m:97-99 (minFunc L-BFGS there, SciPy's L-BFGS-B here) with the objective and
    gradient evaluated on the GPU.  Returns ``(x, f, scipy_result)``; M = x[:P k].reshape(k, P).T etc. (:101-110)."""
    from scipy.optimize import minimize
    ev = TrainingObjective(centered_rest_fluxes, lya_1pzs, rest_noise_variances, k, device)
    try:
        res = minimize(lambda v: ev(v), _f64(initial_x), jac=True, method="L-BFGS-B", options={"maxiter": max_iter})
    finally:
        ev.close()
    return res.x, float(res.fun), res


def matlab_default_rand(n: int) -> np.ndarray:
    """First ``n`` numbers of MATLAB's ``rng('default'); rand`` stream as the library generates them."""
    out = np.empty(int(n))
    _lib.load().gpdla_matlab_default_rand(_dp(out), int(n))
    return out


def voigt(lambdas: Sequence[float], z: float, N: float, num_lines: int = _lib.MAX_LINES) -> np.ndarray:
    """``profile = voigt(lambdas, z, N [, num_lines])`` (voigt.c:253-304): Voigt absorption profile of the
    first ``num_lines`` Lyman-series members for a cloud of column density ``N`` (cm^-2) at redshift ``z``,
    convolved with the 7-pixel instrument profile; returns ``len(lambdas) - 6`` values."""
    lib = _lib.load()
    lam = _f64(lambdas).ravel()
    out = np.empty(max(lam.size - 6, 0))
    _lib.check(lib.gpdla_voigt(_dp(lam), lam.size, float(z), float(N), int(num_lines), _dp(out)))
    return out


def voigt_batch(lambdas, z, N, num_lines: int = _lib.MAX_LINES):
    """Device-resident batched ``voigt``: CUDA float64 tensors ``lambdas`` (P,), ``z`` (S,), ``N`` (S,) ->
    ``(S, P - 6)`` CUDA tensor, on torch's current stream."""
    import torch
    lib = _lib.load()
    assert lambdas.is_cuda and lambdas.dtype == torch.float64 and z.dtype == torch.float64 and N.dtype == torch.float64
    lambdas, z, N = lambdas.contiguous(), z.contiguous(), N.contiguous()
    out = torch.empty((z.numel(), lambdas.numel() - 6), dtype=torch.float64, device=lambdas.device)
    with torch.cuda.device(lambdas.device):
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.gpdla_voigt_batch_device(lambdas.data_ptr(), lambdas.numel(), z.data_ptr(), N.data_ptr(),
                                                z.numel(), int(num_lines), out.data_ptr(), ctypes.c_void_p(stream)))
    return out
