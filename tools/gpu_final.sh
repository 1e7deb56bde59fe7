# final single-GPU validation of a round: GPU tests, smoke, the default bench line
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/final_tests.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final_smoke.log
timeout 900 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/final_bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/final_bench_n1.json'))
print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'], 'roofline', d['roofline']['frac'])
print({k:round(v['ms'],1) for k,v in d['phases'].items()})
print('parity', d['parity']['ok'], d['parity']['S_max_rel_err'])
print('api', d['api_e2e'] and (d['api_e2e']['seconds'], d['api_e2e']['S_max_rel_err']))
print('cpu', d['cpu_baseline'] and (d['cpu_baseline']['kind'], d['cpu_baseline']['cores'], d['cpu_baseline']['sample']))
PY
