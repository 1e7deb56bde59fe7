cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
show() {
python -c "
import json,sys; d=json.load(open('$1')); print('$1', round(d['ms_per_step'],1), round(d['value'],1), 'e2e', round(d['e2e']['ms_per_step'],1), {k:round(v['ms'],1) for k,v in d['phases'].items()}, d['parity'] and (d['parity'].get('S_max_rel_err'), d['parity'].get('ok')))"
}
if [ "$N" = "2" ]; then
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2_tests16.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_tests16.log
fi
timeout 240 $TR --master-port 29513 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/r2_bench_n${N}_e.json 2> gpurun_out/r2_bench_n${N}_e.err; echo "rc=$?"; show gpurun_out/r2_bench_n${N}_e.json
