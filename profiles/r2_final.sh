# final evidence of round 2 on the shipped build: the driver's two commands, every single-GPU bench line of DESIGN §8, the launch list of the default command
mkdir -p gpurun_out/r2f; cd $GRAFT_REPO_ROOT
timeout 600 python bench.py > gpurun_out/r2f/bench_default.json 2> gpurun_out/r2f/err_default.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f/bench_reference.json 2> gpurun_out/r2f/err_reference.txt
B="timeout 400 python bench.py --no-cpu --steps 10 --warmup 5"
$B --config c1 --steps 20 --warmup 10 > gpurun_out/r2f/bench_c1.json 2> gpurun_out/r2f/err_c1.txt
$B --config c2 --storage 2bit > gpurun_out/r2f/bench_c2_2bit.json 2> gpurun_out/r2f/err_c2_2bit.txt
$B --config c5 > gpurun_out/r2f/bench_c5.json 2> gpurun_out/r2f/err_c5.txt
$B --config c5 --storage 2bit > gpurun_out/r2f/bench_c5_2bit.json 2> gpurun_out/r2f/err_c5_2bit.txt
$B --config c3 > gpurun_out/r2f/bench_c3.json 2> gpurun_out/r2f/err_c3.txt
$B --config c3 --storage 2bit > gpurun_out/r2f/bench_c3_2bit.json 2> gpurun_out/r2f/err_c3_2bit.txt
$B --config c2r --no-e2e > gpurun_out/r2f/bench_c2r.json 2> gpurun_out/r2f/err_c2r.txt
$B --config c4 --no-e2e > gpurun_out/r2f/bench_c4.json 2> gpurun_out/r2f/err_c4.txt
$B --config c4rr --no-e2e > gpurun_out/r2f/bench_c4rr.json 2> gpurun_out/r2f/err_c4rr.txt
CMD="python bench.py --no-cpu --no-e2e --steps 2 --warmup 3 --long-seconds 0"
timeout 200 $CMD > gpurun_out/r2f/plain_c2.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2f/launches_c2.csv $CMD > gpurun_out/r2f/ncu_launches.log 2>&1
for f in gpurun_out/r2f/bench_*.json; do echo $f; python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(' ', round(d['ms_per_step'],3),'ms', round(d['value']/1e6,2),'M/s frac',d.get('roofline') and round(d['roofline']['frac'],3), 'e2e', d.get('e2e') and round(d['e2e']['value']/1e6,2), d['config'].get('geometry'))
except Exception as ex: print('  FAILED', ex, open(sys.argv[1]).read()[-300:])
PY
done
