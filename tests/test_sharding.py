"""Host-side multi-GPU logic on CPU: cost-balanced partition and the world_size-2 gloo gather."""
import os
import socket

import numpy as np
import pytest

from gp_dla_detection_b200 import sharding as sh
from gp_dla_detection_b200.api import pad_spectra


def test_partition_by_cost_is_contiguous_and_balanced():
    rng = np.random.default_rng(0)
    costs = rng.integers(200, 1251, size=1000).astype(float)
    for world in (1, 2, 4, 8):
        blocks = sh.partition_by_cost(costs, world)
        assert blocks[0][0] == 0 and blocks[-1][1] == 1000
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
        sums = np.array([costs[s:e].sum() for s, e in blocks])
        assert sums.max() - sums.min() <= 2 * costs.max()
    assert sh.partition_by_cost([5.0], 4)[-1] == (1, 1) or sum(e - s for s, e in sh.partition_by_cost([5.0], 4)) == 1
    assert sh.partition_by_cost([], 2) == [(0, 0), (0, 0)]


def test_costs_agree_between_ragged_and_padded(synthetic_inputs):
    sp = synthetic_inputs["spectra"]
    assert np.array_equal(sh.quasar_costs(sp), sh.quasar_costs(pad_spectra(sp)))


def test_pack_unpack_roundtrip():
    rng = np.random.default_rng(1)
    Q = 7
    res = {n: rng.standard_normal(Q) for n in sh.RECORD_F64}
    res["model_posteriors"] = rng.random((Q, 2))
    res["map_inds"] = np.array([0, 5, 9999, -1, 3, 2, 1], dtype=np.int64)
    res["p_dlas"][3] = np.nan
    back = sh.unpack_records(sh.pack_records(res))
    for k in res:
        assert np.array_equal(back[k], res[k], equal_nan=True), k


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gp_dla_detection_b200 import synthetic as syn
    model = syn.make_model()
    spectra = syn.make_spectra(model, 9, seed=5)

    def fake_compute(sp):   # deterministic stand-in for the CUDA path: results keyed on z_qso
        z = np.asarray(sp["z_qsos"])
        out = {n: z * (i + 1) for i, n in enumerate(sh.RECORD_F64)}
        out["model_posteriors"] = np.stack([z, 1 - z], axis=1)
        out["map_inds"] = (z * 1000).astype(np.int64)
        return out
    res = sh.process_qsos_sharded(model, None, spectra, None, compute=fake_compute)
    ok = (np.allclose(res["p_dlas"], spectra["z_qsos"] * 10) and
          np.array_equal(res["map_inds"], (spectra["z_qsos"] * 1000).astype(np.int64)) and
          res["blocks"][rank][1] > res["blocks"][rank][0])
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_two_rank_gloo_gather():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(60)
    assert sorted(results) == [(0, True), (1, True)]
