cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
show() {
python - "$1" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], 'ms/step', round(d['ms_per_step'],1), 'TF/s', round(d['value'],1), 'e2e', round(d['e2e']['ms_per_step'],1), 'launches', d['gpu_launches'])
print('  ', {k:round(v['ms'],1) for k,v in d['phases'].items()}, 'parity', d['parity'] and (d['parity']['S_max_rel_err'], d['parity']['ok']))
PY
}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 6 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "bench n8 rc=$?"
grep "rank .*ms/step\|e2e timeline" gpurun_out/r2_bench_n8.err | head -16
show gpurun_out/r2_bench_n8.json
MUNK_DIST_THREADS=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 6 --warmup 3 --no-parity > gpurun_out/r2_bench_n8_nothreads.json 2> gpurun_out/r2_bench_n8_nothreads.err; echo "bench n8 nothreads rc=$?"
show gpurun_out/r2_bench_n8_nothreads.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 6 --warmup 3 --no-parity > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err; echo "bench n4 rc=$?"
show gpurun_out/r2_bench_n4.json
