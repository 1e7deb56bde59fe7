"""Multi-GPU check of munk_b200.dist (run with torchrun under `gpurun --gpus N`).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/dist_check.py [n1 n2 k]

Every rank runs the distributed pipeline; rank 0 also runs the single-GPU pipeline and compares
the gathered score matrix with it, then prints timings (device time, max over ranks)."""
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
import synth as synthetic
from munk_b200.device import Device

n1 = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
n2 = int(sys.argv[2]) if len(sys.argv) > 2 else 2200
k = int(sys.argv[3]) if len(sys.argv) > 3 else 100
G1, G2, homs = synthetic.make_case(n1, n2, 6, max(2 * k, 50), seeds=(1, 2))
inp = pipeline.graph_inputs(G1, G2, homs, k)
s_pos, t_pos = inp["s_pos"], inp["t_pos"]
homolog_idxs = [(s_pos[a], t_pos[b]) for a, b in homs]
dev = Device(local)


def run():
    return mdist.dist_embed_scores(dev, n1, inp["edges1"], n2, inp["edges2"], inp["landmark_idxs"],
                                   rank, world, homolog_idxs=homolog_idxs)


out = run()
torch.cuda.synchronize()
# gather S rows on rank 0
S_full = None
if world > 1:
    sizes = [len(mdist.ColumnCyclic(n1, world).rows_of(r)) for r in range(world)]
    mx = max(sizes)
    pad = torch.zeros((mx, n2), dtype=torch.float64, device=dev.torch_device)
    pad[:len(out["rows"])] = out["S"][:, :n2]
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    if rank == 0:
        S_full = torch.empty((n1, n2), dtype=torch.float64, device=dev.torch_device)
        for r in range(world):
            rows = mdist.ColumnCyclic(n1, world).rows_of(r)
            S_full[torch.as_tensor(rows, device=dev.torch_device)] = parts[r][:len(rows)]
else:
    S_full = out["S"][:, :n2]

ok = True
if rank == 0:
    emb = pipeline.embed_device(n1, inp["edges1"], n2, inp["edges2"], inp["landmark_idxs"], dev=dev)
    S_one = emb.scores()[:, :n2]
    err = float((S_full - S_one).abs().max() / S_one.abs().max())
    from munk_b200.scores_plan import plan_separate_scores
    plan = plan_separate_scores(n1, n2, inp["landmark_idxs"], homolog_idxs)
    m = plan.means(dev, emb.scores()).cpu().numpy()
    st = out["stats"]
    err_stats = max(abs(st[0] - m[0]) / abs(m[0]), abs(st[1] - m[1]) / abs(m[1]))
    ok = err < 1e-10 and err_stats < 1e-9
    print("world %d  n1 %d n2 %d k %d : score rel err vs 1 GPU %.2e, stats rel err %.2e -> %s" % (
        world, n1, n2, k, err, err_stats, "OK" if ok else "FAIL"), flush=True)
    del emb, S_one

# host shards through a ShardSink: bit-identical to the device results of the same call
sink = mdist.ShardSink(dev, n1, n2, rank, world)
o2 = mdist.dist_embed_scores(dev, n1, inp["edges1"], n2, inp["edges2"], inp["landmark_idxs"], rank, world,
                             sink=sink)
sink.wait()
torch.cuda.synchronize()
same = (torch.equal(sink.C1, o2["C1"].cpu()) and torch.equal(sink.S, o2["S"][:, :n2].cpu())
        and torch.equal(sink.C2hat, o2["P"][:, sink.col0:sink.col1].t().cpu())
        and sink.rows == o2["rows"])
flag = torch.tensor([0 if same else 1], device=dev.torch_device)
if world > 1:
    dist.all_reduce(flag, op=dist.ReduceOp.MAX)
if rank == 0:
    print("ShardSink host shards equal the device results on every rank: %s" % ("OK" if int(flag) == 0 else "FAIL"),
          flush=True)
ok = ok and int(flag) == 0
del o2

# timing
for _ in range(2):
    run()
ts = []
for _ in range(3):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev.torch_device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ts.append(float(t))
if rank == 0:
    print("world %d pipeline ms (max over ranks): %s" % (world, ["%.1f" % x for x in ts]), flush=True)
if world > 1:
    dist.destroy_process_group()
sys.exit(0 if ok else 1)
