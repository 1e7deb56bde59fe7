cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
show() {
python -c "
import json,sys; d=json.load(open('$1')); print('$1', round(d['ms_per_step'],1), round(d['value'],1), 'e2e', round(d['e2e']['ms_per_step'],1), {k:round(v['ms'],1) for k,v in d['phases'].items()}, d['parity'])"
}
if [ "$2" != "c5only" ]; then
timeout 300 $TR --master-port 29513 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/r2_bench_n${N}_d.json 2> gpurun_out/r2_bench_n${N}_d.err; echo "rc=$?"; show gpurun_out/r2_bench_n${N}_d.json
fi
timeout 420 $TR --master-port 29514 bench.py --gpus $N --config C5 --steps 3 --warmup 3 > gpurun_out/r2_bench_c5_n${N}.json 2> gpurun_out/r2_bench_c5_n${N}.err; echo "rc=$?"; show gpurun_out/r2_bench_c5_n${N}.json; tail -3 gpurun_out/r2_bench_c5_n${N}.err | cut -c1-400
