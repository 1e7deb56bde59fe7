"""Per-column device timeline of the distributed source factorisation (MUNK_TRACE_DIST=1), rank 0 and
the last rank; run under torchrun.  usage: torchrun ... tools/dist_trace.py [C3]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from munk_b200 import dist as mdist, pipeline
from munk_b200.device import Device
import synth

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
n1, n2, m, H, k = synth.CONFIGS[cfg]
G1, G2, homs = synth.make_case(n1, n2, m, H)
inp = pipeline.graph_inputs(G1, G2, homs, k)
del G1, G2
dev = Device(local)
run = lambda: mdist.dist_embed_scores(dev, n1, inp["edges1"], n2, inp["edges2"], inp["landmark_idxs"], rank, world)
for _ in range(3):
    run()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
if rank in (0, world - 1):
    os.environ["MUNK_TRACE_DIST"] = "1"
run()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
