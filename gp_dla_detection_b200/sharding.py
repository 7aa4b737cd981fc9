"""Quasar sharding across the GPUs of one box.

The reference parallelises by hand: ``test_ind`` selects a slice of quasars per SLURM job and
``CDDF_analysis/sbatch_reunion.py:13-63`` concatenates the per-job files.  Here the same split
is a contiguous partition of the catalogue balanced by per-quasar cost (number of pixels in the
modelled window x samples); every rank (one process per GPU) runs the hot path on its block with
no data-path collective, and ONE ``all_gather`` of fixed-width per-quasar records (NCCL over
NVLink on GPUs, gloo in CPU tests) reassembles the catalogue.  ``sample_log_likelihoods_dla``
(80 KB/quasar) stays sharded.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Sequence, Tuple

import numpy as np

RECORD_F64 = ["min_z_dlas", "max_z_dlas", "log_priors_no_dla", "log_priors_dla", "log_likelihoods_no_dla",
              "log_likelihoods_dla", "log_posteriors_no_dla", "log_posteriors_dla", "p_no_dlas", "p_dlas",
              "map_z_dlas", "map_log_nhis"]
RECORD_WIDTH = len(RECORD_F64) + 3   # + model_posteriors (2) + map_inds (1, int64 bit pattern)


def partition_by_cost(costs: Sequence[float], world_size: int) -> List[Tuple[int, int]]:
    """Contiguous blocks ``[start, end)`` whose summed cost is as equal as prefix sums allow."""
    costs = np.asarray(costs, dtype=np.float64)
    Q = costs.size
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    csum = np.concatenate([[0.0], np.cumsum(costs)])
    total = csum[-1]
    bounds = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        j = int(np.searchsorted(csum, target, side="left"))
        if j > 0 and abs(csum[j - 1] - target) <= abs(csum[min(j, Q)] - target):
            j -= 1
        bounds.append(min(max(j, bounds[-1]), Q))
    bounds.append(Q)
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


def quasar_costs(spectra: Dict, min_lambda: float = 911.75, max_lambda: float = 1215.75) -> np.ndarray:
    """Pixels inside the modelled rest-frame window per quasar (cost of one quasar ~ n_u x samples)."""
    if "lengths" in spectra:
        W, L, z = spectra["wavelengths"], spectra["lengths"], spectra["z_qsos"]
        rest = W / (1.0 + np.asarray(z)[:, None])
        valid = np.arange(W.shape[1])[None, :] < np.asarray(L)[:, None]
        return np.count_nonzero(valid & (rest >= min_lambda) & (rest <= max_lambda), axis=1).astype(np.float64)
    return np.array([np.count_nonzero((w / (1 + z) >= min_lambda) & (w / (1 + z) <= max_lambda))
                     for w, z in zip(spectra["all_wavelengths"], spectra["z_qsos"])], dtype=np.float64)


def slice_spectra(spectra: Dict, start: int, end: int) -> Dict:
    if "lengths" in spectra:
        return {k: v[start:end] for k, v in spectra.items()}
    return {k: (v[start:end] if k in ("all_wavelengths", "all_flux", "all_noise_variance", "all_pixel_mask", "z_qsos")
                else v) for k, v in spectra.items()}


def pack_records(res: Dict[str, np.ndarray]) -> np.ndarray:
    """Per-quasar results -> ``[Q_local x RECORD_WIDTH]`` float64 (map_inds carried as its bit pattern)."""
    Q = len(res["p_dlas"])
    rec = np.empty((Q, RECORD_WIDTH), dtype=np.float64)
    for i, n in enumerate(RECORD_F64):
        rec[:, i] = res[n]
    rec[:, len(RECORD_F64):len(RECORD_F64) + 2] = res["model_posteriors"]
    rec[:, -1] = np.asarray(res["map_inds"], dtype=np.int64).view(np.float64)
    return rec


def unpack_records(rec: np.ndarray) -> Dict[str, np.ndarray]:
    out = {n: rec[:, i].copy() for i, n in enumerate(RECORD_F64)}
    out["model_posteriors"] = rec[:, len(RECORD_F64):len(RECORD_F64) + 2].copy()
    out["map_inds"] = rec[:, -1].copy().view(np.int64)
    return out


def gather_records(local: np.ndarray, blocks: List[Tuple[int, int]], device=None) -> np.ndarray:
    """One all_gather of the padded record blocks; returns the catalogue-ordered ``[Q x RECORD_WIDTH]``."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    qmax = max(e - s for s, e in blocks)
    buf = torch.zeros((max(qmax, 1), RECORD_WIDTH), dtype=torch.float64, device=device)
    if local.shape[0]:
        buf[:local.shape[0]] = torch.from_numpy(local).to(buf.device)
    out = torch.empty((world * buf.shape[0], RECORD_WIDTH), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(out, buf)
    out = out.cpu().numpy().reshape(world, buf.shape[0], RECORD_WIDTH)
    return np.concatenate([out[r, :e - s] for r, (s, e) in enumerate(blocks)], axis=0)


def process_qsos_sharded(model: Dict, samples: Dict, spectra: Dict, prior: Dict, params=None,
                         compute: Callable[[Dict], Dict[str, np.ndarray]] = None, device=None) -> Dict:
    """Every rank calls this with the FULL catalogue description; rank r processes block r and all
    ranks return the gathered per-quasar results.  ``compute`` defaults to the CUDA path on
    ``cuda:LOCAL_RANK``; tests inject a CPU stand-in to exercise the partition/gather logic."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    blocks = partition_by_cost(quasar_costs(spectra), world)
    s, e = blocks[rank]
    mine = slice_spectra(spectra, s, e)
    if compute is None:
        import os
        from .api import DLAProcessor
        from .params import DEFAULT
        proc = DLAProcessor(model, samples, prior, params or DEFAULT, device=int(os.environ.get("LOCAL_RANK", 0)))
        compute = lambda sp: proc.process(sp, return_sample_log_likelihoods=False)
    if e > s:
        rec = pack_records(compute(mine))
    else:
        rec = np.zeros((0, RECORD_WIDTH))
    full = gather_records(rec, blocks, device=device)
    out = unpack_records(full)
    out["blocks"] = blocks
    return out
