set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests1.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests1.log
tail -5 gpurun_out/r2_tests1.log
python tools/factor_probe.py 20000 --trace > gpurun_out/r2_probe_default.log 2>&1
MUNK_PANEL_BIG=512 python tools/factor_probe.py 20000 > gpurun_out/r2_probe_nobig.log 2>&1
MUNK_PANEL_BIG=512 MUNK_PANEL_SWITCH=0 python tools/factor_probe.py 20000 > gpurun_out/r2_probe_512only.log 2>&1
MUNK_PANEL_SWITCH=8192 python tools/factor_probe.py 20000 > gpurun_out/r2_probe_sw8192.log 2>&1
MUNK_TRTRI_FORK_DEPTH=0 python tools/factor_probe.py 20000 > gpurun_out/r2_probe_nofork.log 2>&1
python tools/factor_probe.py 16000 > gpurun_out/r2_probe_16000.log 2>&1
tail -n 3 gpurun_out/r2_probe_*.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2_bench1.json
