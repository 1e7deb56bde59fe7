set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export MUNK_GRAPH_DEBUG=1
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests3.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests3.log
tail -25 gpurun_out/r2_tests3.log
python tools/host_probe.py 20000 > gpurun_out/r2_host_probe3.log 2>&1; tail -8 gpurun_out/r2_host_probe3.log
python bench.py --steps 5 --warmup 3 --no-cpu --no-api > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err; echo "bench rc=$?"
tail -5 gpurun_out/r2_bench3.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench3.json'))
print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['gpu_launches'])
print({k:round(v['ms'],2) for k,v in d['phases'].items()})
print(d['parity'])
PY
