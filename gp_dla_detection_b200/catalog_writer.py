"""ASCII catalogue of the DLA search results -- the step after the hot path
(``generate_ascii_catalog.m``).  Pure host-side formatting of the arrays
``process_qsos`` returns; byte-for-byte the reference's ``fprintf`` formats."""
from __future__ import annotations

import re
from typing import Dict, Optional, Sequence

import numpy as np


def _exp3(x: float) -> str:
    """``regexprep(sprintf('%0.5e', x), 'e([+-])(\\d\\d)$', 'e$10$2')`` (generate_ascii_catalog.m:66-69):
    force a three-digit exponent."""
    s = "%0.5e" % x if np.isfinite(x) else ("NaN" if np.isnan(x) else ("Inf" if x > 0 else "-Inf"))
    return re.sub(r"e([+-])(\d\d)$", r"e\g<1>0\2", s)


_SPEC = re.compile(r"%(?P<flags>[-0 +#]*)(?P<width>\d*)(?:\.(?P<prec>\d+))?(?P<conv>[feEgG])")


def _mfmt(spec: str, x: float) -> str:
    """One ``fprintf`` float conversion as MATLAB prints it: finite values like C, non-finite ones as ``NaN`` / ``Inf`` /
    ``-Inf`` padded with blanks to the field width (MATLAB ignores the zero flag there; C and Python print ``nan``)."""
    if np.isfinite(x):
        return spec % x
    m = _SPEC.fullmatch(spec)
    text = "NaN" if np.isnan(x) else ("Inf" if x > 0 else "-Inf")
    width = int(m.group("width") or 0)
    return text.ljust(width) if "-" in m.group("flags") else text.rjust(width)


def write_dla_samples(path: str, offset_samples: Sequence[float], log_nhi_samples: Sequence[float]) -> None:
    """``<test_set_name>_dla_samples.dat`` (generate_ascii_catalog.m:9-20)."""
    with open(path, "w") as f:
        for o, n in zip(offset_samples, log_nhi_samples):
            f.write("%06f %09f\n" % (o, n))


def write_results(path: str, results: Dict[str, np.ndarray], thing_ids: Sequence[int],
                  offset_samples: Optional[Sequence[float]] = None,
                  log_nhi_samples: Optional[Sequence[float]] = None) -> None:
    """``<test_set_name>_results.dat`` (generate_ascii_catalog.m:49-81): one line per searched quasar with
    z range, log priors, log likelihoods, model posteriors and the MAP (z_DLA, log N_HI).  The reference
    passes one argument to ``'%09i %-18s '`` so only the THING_ID is printed; that is kept.  MAP values come
    from ``results['map_z_dlas'/'map_log_nhis']`` (already computed on the GPU) or, if absent, from
    ``sample_log_likelihoods_dla`` and the sample arrays like the reference does (:73-80)."""
    Q = len(results["p_dlas"])
    if "map_z_dlas" in results:
        map_z, map_n = results["map_z_dlas"], results["map_log_nhis"]
    else:
        sll = results["sample_log_likelihoods_dla"]
        idx = np.array([np.nanargmax(sll[i]) for i in range(Q)])
        map_z = results["min_z_dlas"] + (results["max_z_dlas"] - results["min_z_dlas"]) * np.asarray(offset_samples)[idx]
        map_n = np.asarray(log_nhi_samples)[idx]
    mp = results["model_posteriors"]
    with open(path, "w") as f:
        for i in range(Q):
            f.write("%09i " % int(thing_ids[i]))
            f.write(" ".join([
                _mfmt("%06.4f", results["min_z_dlas"][i]), _mfmt("%06.4f", results["max_z_dlas"][i]),
                _mfmt("%8.5f", results["log_priors_no_dla"][i]), _mfmt("%8.5f", results["log_priors_dla"][i]),
                _mfmt("%12.5e", results["log_likelihoods_no_dla"][i]), _mfmt("%12.5e", results["log_likelihoods_dla"][i]),
                _exp3(mp[i, 0]), _exp3(mp[i, 1])]) + " ")
            f.write("%s %s\n" % (_mfmt("%06.4f", map_z[i]), _mfmt("%07.4f", map_n[i])))


def _item(x):
    """numpy scalar -> Python scalar (``.item()`` in the reference), so that ``json`` can serialise it."""
    return x.item() if hasattr(x, "item") else x


def write_json_catalogue(path: str, results: Dict[str, np.ndarray], info: Dict[str, Sequence], sub_dla: bool = True):
    """``predictions_multi_DLAs.json`` (CDDF_analysis/qso_loader.py:1927-2033, ``generate_json_catalogue``), the
    Parks-et-al.-style catalogue of the multi-DLA run: one record per quasar with ``p_dla``, ``p_no_dla``, the
    largest model posterior, the number of DLAs of the most probable model and their MAP ``(z_dla, log_nhi)``.

    ``results`` is what ``process_qsos_multiple_dlas_meanflux`` returns (``model_posteriors`` with columns
    [no DLA, sub-DLA, 1..max DLAs], ``MAP_z_dlas`` / ``MAP_log_nhis`` ``[Q, max_dlas, max_dlas]``, ``p_dlas``,
    ``p_no_dlas``, ``min_z_dlas``, ``max_z_dlas``); ``info`` carries the per-quasar catalogue columns ``thing_ids,
    z_qsos, snrs, ras, decs, plates, mjds, fiber_ids``.  As in the reference, a quasar whose most probable model is
    the null or the sub-DLA model reports ``p_no_dla`` as its ``max_model_posterior`` and zero DLAs (:1970-1979).
    Returns the list that was written."""
    import json
    mp = np.asarray(results["model_posteriors"])
    model_index = np.array([int(np.argmax(r)) if not np.all(np.isnan(r)) else 0 for r in mp])
    max_mp = np.array([r[i] for r, i in zip(mp, model_index)], dtype=np.float64)
    off = 1 if sub_dla else 0
    if sub_dla:
        inds = model_index < 1 + off
        max_mp[inds] = np.asarray(results["p_no_dlas"])[inds]
    num_dlas = np.maximum(model_index - off, 0)
    out = []
    for i, thing_id in enumerate(info["thing_ids"]):
        n = int(num_dlas[i])
        spec = {
            "p_dla": _item(results["p_dlas"][i]), "p_no_dla": _item(results["p_no_dlas"][i]),
            "max_model_posterior": float(max_mp[i]), "num_dlas": n,
            "min_z_dla": _item(results["min_z_dlas"][i]), "max_z_dla": _item(results["max_z_dlas"][i]),
            "snr": _item(info["snrs"][i]), "ra": _item(info["ras"][i]), "dec": _item(info["decs"][i]),
            "plate": _item(info["plates"][i]), "mjd": _item(info["mjds"][i]), "fiber_id": _item(info["fiber_ids"][i]),
            "thing_id": _item(thing_id), "z_qso": _item(info["z_qsos"][i]),
        }
        spec["dlas"] = [{"log_nhi": float(results["MAP_log_nhis"][i, n - 1, j]),
                         "z_dla": float(results["MAP_z_dlas"][i, n - 1, j])} for j in range(n)]
        out.append(spec)
    with open(path, "w") as f:
        json.dump(out, f, indent=2)
    return out


def write_sub_dla_catalogue(path: str, results: Dict[str, np.ndarray], info: Dict[str, Sequence]):
    """``predictions_sub_DLA_candidates.json`` (CDDF_analysis/qso_loader.py:2035-2090): the quasars whose most
    probable model is the sub-DLA model (column 1 of ``model_posteriors``), with that posterior as ``p_sub_dla``."""
    import json
    mp = np.asarray(results["model_posteriors"])
    out = []
    for i in range(mp.shape[0]):
        if np.all(np.isnan(mp[i])) or int(np.argmax(mp[i])) != 1:
            continue
        out.append({"p_sub_dla": float(mp[i, 1]), "ra": _item(info["ras"][i]), "snr": _item(info["snrs"][i]),
                    "dec": _item(info["decs"][i]), "plate": _item(info["plates"][i]), "mjd": _item(info["mjds"][i]),
                    "fiber_id": _item(info["fiber_ids"][i]), "thing_id": _item(info["thing_ids"][i]),
                    "z_qso": _item(info["z_qsos"][i])})
    with open(path, "w") as f:
        json.dump(out, f, indent=2)
    return out


# ---------------------------------------------------------------------------------------------------------------
# MATLAB-layout result arrays: the variables process_qsos.m:236-250 / ...meanflux.m:498-523 save.

SINGLE_VARIABLES = ["min_z_dlas", "max_z_dlas", "log_priors_no_dla", "log_priors_dla", "log_likelihoods_no_dla",
                    "sample_log_likelihoods_dla", "log_likelihoods_dla", "log_posteriors_no_dla", "log_posteriors_dla",
                    "model_posteriors", "p_no_dlas", "p_dlas"]
MULTI_VARIABLES = ["min_z_dlas", "max_z_dlas", "sample_log_likelihoods_dla", "base_sample_inds", "log_priors_no_dla",
                   "log_priors_dla", "log_priors_lls", "log_likelihoods_no_dla", "MAP_z_dlas", "MAP_log_nhis",
                   "log_likelihoods_dla", "log_likelihoods_lls", "log_posteriors_no_dla", "log_posteriors_dla",
                   "log_posteriors_lls", "model_posteriors", "p_no_dlas", "p_dlas", "p_lls", "all_exceptions",
                   "sample_log_likelihoods_lls"]


def matlab_arrays(results: Dict[str, np.ndarray], run_info: Optional[Dict] = None, multi: bool = False,
                  small_file: bool = False) -> Dict[str, np.ndarray]:
    """The result variables in the shapes MATLAB holds them in (process_qsos.m:74-82, ...meanflux.m:109-139): per-quasar
    vectors as ``(Q, 1)`` columns, ``sample_log_likelihoods_dla`` ``(Q, S)`` / ``(Q, S, max_dlas)``, ``base_sample_inds``
    ``(Q, S, max_dlas - 1)`` uint32 and 1-based with 0 where a level was never reached, ``MAP_*`` ``(Q, max_dlas,
    max_dlas)``, ``all_exceptions`` 1 for spectra without a usable pixel and NaN otherwise (:139, :232).  ``run_info``
    adds the scalars / strings the scripts save next to them (``training_release``, ``test_ind``, ``num_lines``, ...).
    ``small_file`` drops the per-sample arrays as ``save2mat73(small_file=True)`` does (sbatch_reunion.py:78-81)."""
    names = MULTI_VARIABLES if multi else SINGLE_VARIABLES
    Q = len(results["p_dlas"])
    out: Dict[str, np.ndarray] = {}
    for n in names:
        if small_file and ("sample_log_likelihoods" in n or n == "base_sample_inds"):
            continue
        if n == "all_exceptions":
            out[n] = np.where(np.isnan(np.asarray(results["min_z_dlas"], dtype=np.float64)), 1.0, np.nan).reshape(Q, 1)
            continue
        if n not in results:
            continue
        a = np.asarray(results[n])
        if n == "base_sample_inds":
            ll = np.asarray(results["log_likelihoods_dla"])                       # (Q, max_dlas)
            reached = np.isfinite(ll[:, :a.shape[2]])                              # level r + 1 finished -> row r was drawn
            a = np.where(reached[:, None, :], a.astype(np.int64) + 1, 0).astype(np.uint32)
        elif a.ndim == 1:
            a = a.reshape(Q, 1).astype(np.float64)
        out[n] = a
    for k, v in (run_info or {}).items():
        out[k] = v
    return out


def h5py_view(arrays: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """What ``h5py.File(processed_file)[name][()]`` returns for a MATLAB v7.3 file holding ``arrays``: every array with
    its axes reversed (MATLAB is column-major), i.e. ``(1, Q)`` vectors, ``model_posteriors`` ``(2, Q)``,
    ``sample_log_likelihoods_dla`` ``(max_dlas, S, Q)`` -- the layout CDDF_analysis/qso_loader.py:84-110 indexes
    (``f['p_dlas'][0, :]``, ``f['model_posteriors'][()].T``) and sbatch_reunion.py:65-85 undoes with ``np.transpose``."""
    return {k: (np.transpose(v) if isinstance(v, np.ndarray) and v.ndim >= 2 else v) for k, v in arrays.items()}


def write_processed_mat(path: str, results: Dict[str, np.ndarray], run_info: Optional[Dict] = None, multi: bool = False,
                        small_file: bool = False) -> Dict[str, np.ndarray]:
    """``processed_qsos_<test_set_name>.mat`` with the reference's variable names and MATLAB shapes
    (process_qsos.m:236-250; multi-DLA: ...meanflux.m:498-523).  Written with ``scipy.io.savemat`` (MAT v5,
    compressed): MATLAB / Octave ``load`` it directly and ``save(..., '-v7.3')`` turns it into the HDF5 container the
    reference's Python readers open (h5py / hdf5storage are not in this image; ``h5py_view`` gives the arrays as those
    readers would see them).  Returns the dict that was written."""
    from scipy.io import savemat
    arrays = matlab_arrays(results, run_info, multi, small_file)
    savemat(path, arrays, do_compression=True, oned_as="column")
    return arrays
