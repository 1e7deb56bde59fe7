cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 120 python tools/factor_probe.py 2048 > gpurun_out/r2_probe10_2048.log 2>&1; echo "probe 2048 rc=$?"; tail -2 gpurun_out/r2_probe10_2048.log
timeout 180 python tools/factor_probe.py 20000 --trace > gpurun_out/r2_probe10_20000.log 2>&1; echo "probe 20000 rc=$?"; head -45 gpurun_out/r2_probe10_20000.log
MUNK_FUSED_DIAG=0 timeout 180 python tools/factor_probe.py 20000 > gpurun_out/r2_probe10_20000_nofused.log 2>&1; tail -1 gpurun_out/r2_probe10_20000_nofused.log
MUNK_STRIP_ROWS=0 timeout 180 python tools/factor_probe.py 20000 > gpurun_out/r2_probe10_20000_nostrip.log 2>&1; tail -1 gpurun_out/r2_probe10_20000_nostrip.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests10.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r2_tests10.log
show() {
python - "$1" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], round(d['ms_per_step'],1), round(d['value'],2), 'e2e', round(d['e2e']['ms_per_step'],1), d['gpu_launches'])
print('  ', {k:round(v['ms'],1) for k,v in d['phases'].items()})
PY
}
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-api --no-parity > gpurun_out/r2_bench10.json 2> gpurun_out/r2_bench10.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench10.err; show gpurun_out/r2_bench10.json
