# usage: bash tools/gpu_sanitize.sh memcheck|racecheck|synccheck     (ONE tool per gpurun call)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
tool=$1
python tools/sanitize_target.py > gpurun_out/r2_sanitize_plain_$tool.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2_sanitize_plain_$tool.log; exit 1; }
timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_target.py > gpurun_out/r2_sanitize_$tool.log 2>&1
echo "compute-sanitizer $tool rc=$?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|all checks passed|Error|hazard" gpurun_out/r2_sanitize_$tool.log | head -20
