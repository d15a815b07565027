# A/B on one box: the build of the last commit (profiles/tmp_wt) against the working tree, same bench command
mkdir -p gpurun_out/ab2; cd $GRAFT_REPO_ROOT
A="--no-cpu --no-e2e --steps 10 --warmup 5 --long-seconds 0"
run() { (cd $1 && timeout 120 python bench.py $A $3 2>/dev/null | tail -1) > gpurun_out/ab2/$2.json; python - gpurun_out/ab2/$2.json $2 <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3), d['config']['geometry']['block'], d['config']['geometry']['tile_stages'])
except Exception as ex: print(sys.argv[2], 'FAILED', ex)
PY
}
run profiles/tmp_wt old_c3_1 "--config c3"
run . new_c3_1 "--config c3"
run profiles/tmp_wt old_c3_2 "--config c3"
run . new_c3_2 "--config c3"
run profiles/tmp_wt old_c2_1 "--config c2"
run . new_c2_1 "--config c2"
