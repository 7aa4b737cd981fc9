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
            f.write("%06.4f %06.4f %8.5f %8.5f %12.5e %12.5e %s %s " % (
                results["min_z_dlas"][i], results["max_z_dlas"][i], results["log_priors_no_dla"][i],
                results["log_priors_dla"][i], results["log_likelihoods_no_dla"][i],
                results["log_likelihoods_dla"][i], _exp3(mp[i, 0]), _exp3(mp[i, 1])))
            f.write("%06.4f %07.4f\n" % (map_z[i], map_n[i]))
