cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 tools/dist_trace.py C3 > gpurun_out/r2_dist_trace_n8.log 2>&1; echo "trace rc=$?"
grep -A42 "dist potrf n=20000 rank 0" gpurun_out/r2_dist_trace_n8.log | head -45
show() {
python - "$1" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], 'ms/step', round(d['ms_per_step'],1), 'TF/s', round(d['value'],1), 'e2e', round(d['e2e']['ms_per_step'],1), 'launches', d['gpu_launches'])
print('  ', {k:round(v['ms'],1) for k,v in d['phases'].items()})
PY
}
MUNK_SPLITK_MIN_IT=12 timeout 600 $TR --master-port 29522 bench.py --gpus 8 --steps 6 --warmup 3 --no-parity > gpurun_out/r2_bench_n8_split12.json 2> gpurun_out/r2_bench_n8_split12.err; echo "bench rc=$?"
show gpurun_out/r2_bench_n8_split12.json
timeout 900 $TR --master-port 29523 tools/perm_bench.py --perms 1000 > gpurun_out/r2_perm_dense_n8_1000.json 2> gpurun_out/r2_perm_dense_n8_1000.err; echo "perm rc=$?"
tail -c 1500 gpurun_out/r2_perm_dense_n8_1000.json
