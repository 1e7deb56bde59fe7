cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2_multi_tests.log 2>&1; echo "multi tests rc=$?"; tail -30 gpurun_out/r2_multi_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py 6000 5000 200 > gpurun_out/r2_dist_check_n2.log 2>&1; echo "dist_check rc=$?"; grep -v "^W\|^\[W" gpurun_out/r2_dist_check_n2.log | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench n2 rc=$?"
grep "rank .*ms/step\|e2e timeline" gpurun_out/r2_bench_n2.err | tail -6
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n2.json'))
print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], d['gpu_launches'])
print({k:round(v['ms'],2) for k,v in d['phases'].items()})
print(d['parity'])
PY
