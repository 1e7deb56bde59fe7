cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests7.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r2_tests7.log
python tools/dual_probe.py > gpurun_out/r2_dual_probe.log 2>&1; cat gpurun_out/r2_dual_probe.log | tail -6
show() {
python - "$1" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], round(d['ms_per_step'],1), round(d['value'],2), 'e2e', round(d['e2e']['ms_per_step'],1), d['gpu_launches'])
print('  ', {k:round(v['ms'],1) for k,v in d['phases'].items()})
print('  ', d.get('api_e2e'))
PY
}
MUNK_GRAPH_DEBUG=1 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_bench7.json 2> gpurun_out/r2_bench7.err; echo "bench rc=$?"; tail -4 gpurun_out/r2_bench7.err
show gpurun_out/r2_bench7.json
MUNK_PAIR=0 python bench.py --steps 5 --warmup 3 --no-cpu --no-api --no-parity > gpurun_out/r2_bench7_nopair.json 2> gpurun_out/r2_bench7_nopair.err; show gpurun_out/r2_bench7_nopair.json
MUNK_PAIR_LEAD_PCT=100 python bench.py --steps 5 --warmup 3 --no-cpu --no-api --no-parity > gpurun_out/r2_bench7_lead100.json 2> gpurun_out/r2_bench7_lead100.err; show gpurun_out/r2_bench7_lead100.json
MUNK_PAIR_LEAD_PCT=200 python bench.py --steps 5 --warmup 3 --no-cpu --no-api --no-parity > gpurun_out/r2_bench7_lead200.json 2> gpurun_out/r2_bench7_lead200.err; show gpurun_out/r2_bench7_lead200.json
python tools/cli_wall.py C3 > gpurun_out/r2_cli_wall_c3.log 2>&1; tail -22 gpurun_out/r2_cli_wall_c3.log
