cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/profile_target_diag.py > gpurun_out/r2_profile_diag_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"potrf_diag" -c 3 \
    -o gpurun_out/r2_prof_diag python tools/profile_target_diag.py > gpurun_out/r2_ncu_diag.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2_ncu_diag.log; cat gpurun_out/r2_profile_diag_plain.log
