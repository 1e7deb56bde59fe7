cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_kernels.py -x -q > gpurun_out/r2_tests12.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2_tests12.log
timeout 300 $TR --master-port 29511 tools/dist_check.py 6000 5000 200 > gpurun_out/r2_dist_check_n2_b.log 2>&1; echo "dist_check rc=$?"; grep "world\|Shard" gpurun_out/r2_dist_check_n2_b.log
timeout 300 $TR --master-port 29512 bench.py --gpus 2 --config C1 --steps 3 --warmup 3 --parity-properties > gpurun_out/r2_bench_n2_c1.json 2> gpurun_out/r2_bench_n2_c1.err; echo "bench c1 rc=$?"; tail -2 gpurun_out/r2_bench_n2_c1.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n2_c1.json')); print(d['ms_per_step'], d['parity'])"
timeout 600 $TR --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2_b.json 2> gpurun_out/r2_bench_n2_b.err; echo "bench n2 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n2_b.json')); print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], {k:round(v['ms'],1) for k,v in d['phases'].items()}, d['parity']['S_max_rel_err'], d['parity']['ok'])"
