"""BASELINE config 4: homology-scores permutation test on the C3 graphs, permutations sharded
round-robin over the ranks (munk_b200/permtest.py).  Run with torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29531 tools/perm_bench.py --perms 64 [--config C3]

Prints one JSON line: permutations/s over all ranks (device time, max over ranks), the time per
permutation on one GPU, and the statistic of the real pair.  1000 permutations = 1000/value s.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--perms", type=int, default=64)
ap.add_argument("--config", default="C3")
ap.add_argument("--rank-k", dest="rank_k", action="store_true",
                help="opt-in fast path: statistic from the factored form, S never formed")
args = ap.parse_args()

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

from munk_b200 import permtest, pipeline
import synth as synthetic
from munk_b200.device import Device

n1, n2, m, H, k = synthetic.CONFIGS[args.config]
G1, G2, homs = synthetic.make_case(n1, n2, m, H)
inp = pipeline.graph_inputs(G1, G2, homs, k)
homolog_idxs = [(inp["s_pos"][a], inp["t_pos"][b]) for a, b in homs]
dev = Device(local)

t0 = time.time()
pair = permtest.FactoredPair(dev, n1, inp["edges1"], n2, inp["edges2"])
real = pair.statistic(homolog_idxs, k).cpu().numpy()
torch.cuda.synchronize()
t_setup = time.time() - t0

mine = permtest.shard(args.perms, rank, world)
pair.statistic(permtest.permuted_homolog_idxs(homolog_idxs, 10 ** 6), k)      # warm-up
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
if args.rank_k:
    pending = {p: pair.statistic_rank_k(permtest.permuted_homolog_idxs(homolog_idxs, p), k) for p in mine}
else:
    pending = {p: pair.statistic(permtest.permuted_homolog_idxs(homolog_idxs, p), k) for p in mine}
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev.torch_device)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
local_res = {}
for p, t in pending.items():
    hom_mean, other_mean, diff = t if args.rank_k else t.cpu().numpy()[:3]
    local_res[p] = (diff, other_mean, hom_mean)
table = permtest.gather_results(local_res, args.perms, rank, world, device=dev.torch_device)
cross = None
if rank == 0 and not args.rank_k:
    # independent formulation (factored form, S never formed) on three of the permutations
    cross = 0.0
    for p in range(min(3, args.perms)):
        a = np.array(pair.statistic_rank_k(permtest.permuted_homolog_idxs(homolog_idxs, p), k))
        b = table[p][[2, 1, 0]]                       # (hom_mean, other_mean, diff)
        cross = max(cross, float(np.max(np.abs(a - b) / np.abs(b))))
if rank == 0:
    diffs = table[:, 0]
    pval, n_less, n_obs = permtest.one_tail_pval(real[2], diffs)
    print(json.dumps({
        "metric": "permutation-test units/s (%s)" % ("rank-k statistic, S never formed" if args.rank_k
                                                      else "embed + dense scores + statistic per permutation"),
        "config": "%s: %d/%d nodes, %d homologs, %d landmarks, 'indices' permutations" % (args.config, n1, n2, H, k),
        "n_gpus": world, "permutations": args.perms, "ms_total": float(ms),
        "value": args.perms / (float(ms) * 1e-3), "unit": "permutations/s",
        "ms_per_permutation_per_gpu": float(ms) / max(1, len(mine)),
        "seconds_for_1000": 1000.0 / (args.perms / (float(ms) * 1e-3)),
        "setup_s_factor_both_graphs": t_setup,
        "real": {"hom_mean": float(real[0]), "other_mean": float(real[1]), "diff": float(real[2])},
        "crosscheck_rank_k_max_rel_diff": cross,
        "pval": pval, "effect_size": float(permtest.effect_size(real[2], diffs)), "errs": int(np.isnan(diffs).sum()),
    }))
if world > 1:
    dist.destroy_process_group()
