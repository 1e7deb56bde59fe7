cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
show() {
python - "$1" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], round(d['ms_per_step'],1), round(d['value'],2), 'e2e', round(d['e2e']['ms_per_step'],1), d['gpu_launches'])
print('  ', {k:round(v['ms'],1) for k,v in d['phases'].items()})
PY
}
python bench.py --steps 5 --warmup 3 --no-cpu --no-api --no-parity > gpurun_out/r2_bench8.json 2> gpurun_out/r2_bench8.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench8.err; show gpurun_out/r2_bench8.json
python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py -m gpu -x -q 2>&1 | tail -3
