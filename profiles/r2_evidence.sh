mkdir -p gpurun_out/r2w; cd $GRAFT_REPO_ROOT
B="timeout 400 python bench.py --no-cpu --steps 10 --warmup 5"
$B --config c1 --steps 20 --warmup 10 > gpurun_out/r2w/bench_c1.json 2> gpurun_out/r2w/err_c1.txt
$B --config c5 > gpurun_out/r2w/bench_c5.json 2> gpurun_out/r2w/err_c5.txt
$B --config c5 --storage 2bit > gpurun_out/r2w/bench_c5_2bit.json 2> gpurun_out/r2w/err_c5_2bit.txt
$B --config c4 --no-e2e > gpurun_out/r2w/bench_c4.json 2> gpurun_out/r2w/err_c4.txt
$B --config c4rr --no-e2e > gpurun_out/r2w/bench_c4rr.json 2> gpurun_out/r2w/err_c4rr.txt
$B --config c3 > gpurun_out/r2w/bench_c3.json 2> gpurun_out/r2w/err_c3.txt
$B --config c3 --storage 2bit > gpurun_out/r2w/bench_c3_2bit.json 2> gpurun_out/r2w/err_c3_2bit.txt
$B --config c3 --n 100000 --p 600000 --model BayesPR --regions 100 --steps 5 --warmup 3 --long-seconds 0 --no-e2e > gpurun_out/r2w/bench_c3_bayespr100.json 2> gpurun_out/r2w/err_c3_pr.txt
$B --config c3 --n 100000 --p 600000 --model BayesPR --regions 100 --steps 5 --warmup 3 --long-seconds 0 --no-e2e --storage 2bit > gpurun_out/r2w/bench_c3_bayespr100_2bit.json 2> gpurun_out/r2w/err_c3_pr2.txt
for f in gpurun_out/r2w/bench_*.json; do echo $f; python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(' ', round(d['ms_per_step'],3),'ms', round(d['value']/1e6,2),'M/s frac',round(d['roofline']['frac'],3), 'e2e', d['e2e'] and round(d['e2e']['value']/1e6,2), d['config']['geometry'])
except Exception as ex: print('  FAILED', ex, open(sys.argv[1]).read()[-300:])
PY
done
