set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests2.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests2.log
tail -15 gpurun_out/r2_tests2.log
python tools/host_probe.py 20000 > gpurun_out/r2_host_probe.log 2>&1; tail -4 gpurun_out/r2_host_probe.log
MUNK_NO_GRAPH=1 python tools/host_probe.py 20000 > gpurun_out/r2_host_probe_nograph.log 2>&1; tail -2 gpurun_out/r2_host_probe_nograph.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench2.json'))
print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['gpu_launches'])
print({k:round(v['ms'],2) for k,v in d['phases'].items()})
print(d['parity']); print(d['api_e2e'])
PY
