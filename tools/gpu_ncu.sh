cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/profile_target.py > gpurun_out/r2_profile_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on \
    -k regex:"dgemm_dmma_tma|trsm_strip|potrf_leaf|fill_reglap|separate_scores_kernel|masked_row_sum" \
    -o gpurun_out/r2_prof python tools/profile_target.py > gpurun_out/r2_ncu.log 2>&1
echo "ncu rc=$?"; tail -5 gpurun_out/r2_ncu.log; ls -la gpurun_out/r2_prof* 2>/dev/null
