"""BASELINE.json configs[2]: the full 162 861-quasar synthetic DR12Q catalogue x 10 000 DLA samples, quasars
sharded over the GPUs of one box (torchrun, one rank per GPU), per-quasar records gathered with one NCCL
all_gather.  Rank 0 spot-checks a few quasars against the oracle and prints one JSON object.

  torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_full_catalog.py [--quasars 162861]
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from gp_dla_detection_b200 import api, sharding, synthetic as syn

ap = argparse.ArgumentParser()
ap.add_argument("--quasars", type=int, default=162861)
ap.add_argument("--check", type=int, default=3)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
Q = args.quasars
blocks = [(Q * r // world, Q * (r + 1) // world) for r in range(world)]   # synthetic spectra have equal cost
s, e = blocks[rank]
model, samples, prior = syn.make_model(), syn.make_samples(10000), syn.make_prior()
t0 = time.perf_counter()
spectra = syn.make_spectra(model, e - s, shard=1000 + rank)
pad = api.pad_spectra(spectra)
t_gen = time.perf_counter() - t0
proc = api.DLAProcessor(model, samples, prior, device=local)
proc.process({k: v[:64] for k, v in pad.items()}, return_sample_log_likelihoods=False)   # warm-up
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
res = proc.process(pad, return_sample_log_likelihoods=False)          # host buffers in, host results out
rec = sharding.pack_records(res)
full = sharding.gather_records(rec, blocks, device=dev) if world > 1 else rec
torch.cuda.synchronize()
dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
if rank == 0:
    from oracle import process_qsos_oracle as O
    allres = sharding.unpack_records(full)
    idx = np.linspace(0, e - s - 1, args.check).astype(int)
    sub = {k: ([v[i] for i in idx] if isinstance(v, list) else v[idx]) for k, v in spectra.items()}
    ref = O.process_qsos(model, samples, sub, prior, engine="c")
    rel = max(abs(allres["log_likelihoods_dla"][i] - ref["log_likelihoods_dla"][j]) / abs(ref["log_likelihoods_dla"][j])
              for j, i in enumerate(idx))
    pd = max(abs(allres["p_dlas"][i] - ref["p_dlas"][j]) for j, i in enumerate(idx))
    same_map = bool(all(allres["map_inds"][i] == ref["map_inds"][j] for j, i in enumerate(idx)))
    print(json.dumps({"workload": "configs[2]: full synthetic DR12Q catalogue", "quasars": Q, "n_gpus": world,
                      "seconds": float(dt.item()), "quasars_per_s": Q / float(dt.item()),
                      "spectra_generation_s_per_rank": t_gen, "records_gathered": int(full.shape[0]),
                      "p_dla_gt_0.9": int(np.sum(allres["p_dlas"] > 0.9)),
                      "spot_check": {"quasars": [int(i) for i in idx], "max_rel_err_log_likelihoods_dla": rel,
                                     "max_abs_err_p_dla": pd, "same_map_index": same_map}}))
if world > 1:
    dist.destroy_process_group()
