# per-launch device times of one bench step (direct launches: MUNK_NO_GRAPH=1), for the kernel shares
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export MUNK_NO_GRAPH=1
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-api --no-parity"
$CMD > gpurun_out/r2_launch_plain.log 2> gpurun_out/r2_launch_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 12000 -c 6000 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_launch_ncu.log 2>&1
echo "rc=$?"; wc -l gpurun_out/r2_launches.csv
