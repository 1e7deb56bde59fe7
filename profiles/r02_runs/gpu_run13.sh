cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
show() {
python -c "
import json,sys; d=json.load(open('$1')); print('$1', round(d['ms_per_step'],1), round(d['value'],1), 'e2e', round(d['e2e']['ms_per_step'],1), {k:round(v['ms'],1) for k,v in d['phases'].items()}, d['parity'] and d['parity'].get('ok'))"
}
timeout 600 $TR --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2_c.json 2> gpurun_out/r2_bench_n2_c.err; echo "rc=$?"; show gpurun_out/r2_bench_n2_c.json
MUNK_NCCL_CHAIN_CTAS=0 MUNK_NCCL_BULK_CTAS=0 timeout 600 $TR --master-port 29514 bench.py --gpus 2 --steps 5 --warmup 3 --no-parity > gpurun_out/r2_bench_n2_c_nocap.json 2> gpurun_out/r2_bench_n2_c_nocap.err; echo "rc=$?"; show gpurun_out/r2_bench_n2_c_nocap.json
MUNK_NCCL_CHAIN_CTAS=2 MUNK_NCCL_BULK_CTAS=4 timeout 600 $TR --master-port 29515 bench.py --gpus 2 --steps 5 --warmup 3 --no-parity > gpurun_out/r2_bench_n2_c_cap24.json 2> gpurun_out/r2_bench_n2_c_cap24.err; echo "rc=$?"; show gpurun_out/r2_bench_n2_c_cap24.json
