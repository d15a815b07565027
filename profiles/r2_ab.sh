# A/B on one box: the build of commit 632fc92 (profiles/tmp_wt) against the working tree, same bench command
mkdir -p gpurun_out/ab; cd $GRAFT_REPO_ROOT
A="--no-cpu --no-e2e --steps 10 --warmup 5 --long-seconds 0"
run() { (cd $1 && timeout 300 python bench.py $A $3 2>/dev/null | tail -1) > gpurun_out/ab/$2.json; python - gpurun_out/ab/$2.json $2 <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3), d['config']['geometry']['block'], d['config']['geometry']['tile_stages'])
except Exception as ex: print(sys.argv[2], 'FAILED', ex)
PY
}
for rep in 1 2; do
run profiles/tmp_wt old_c2_$rep "--config c2"
run . new_c2_$rep "--config c2"
run profiles/tmp_wt old_c2_2bit_$rep "--config c2 --storage 2bit"
run . new_c2_2bit_$rep "--config c2 --storage 2bit"
done
run profiles/tmp_wt old_c3_2bit "--config c3 --storage 2bit"
run . new_c3_2bit "--config c3 --storage 2bit"
run profiles/tmp_wt old_c5_2bit "--config c5 --storage 2bit"
run . new_c5_2bit "--config c5 --storage 2bit"
