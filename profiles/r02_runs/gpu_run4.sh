cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" python bench.py --steps 4 --warmup 3 --no-cpu --no-api --no-parity > gpurun_out/r2_b4_$tag.json 2> gpurun_out/r2_b4_$tag.err
  python - "$tag" <<'PY'
import json,sys
d=json.load(open('gpurun_out/r2_b4_%s.json'%sys.argv[1]))
print(sys.argv[1], round(d['ms_per_step'],1), 'e2e', round(d['e2e']['ms_per_step'],1), {k:round(v['ms'],1) for k,v in d['phases'].items()})
PY
}
run default A=1
run src0 MUNK_SRC_PRIORITY=0
run tgthi MUNK_SRC_PRIORITY=0 MUNK_TGT_PRIORITY=-1
run src0_nograph MUNK_SRC_PRIORITY=0 MUNK_NO_GRAPH=1
