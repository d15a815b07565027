# after the shared-tile change (C5-like int8 shapes): full GPU suite, then the affected bench lines; every command under a short timeout
mkdir -p gpurun_out/r2g; cd $GRAFT_REPO_ROOT
timeout 60 python -m pytest tests/test_gpu_parity.py -x -q -k "many_rows" 2>&1 | tail -2 || exit 1
timeout 280 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
B="timeout 150 python bench.py --no-cpu --steps 10 --warmup 5"
$B --config c5 > gpurun_out/r2g/bench_c5.json 2> gpurun_out/r2g/err_c5.txt
$B --config c3 > gpurun_out/r2g/bench_c3.json 2> gpurun_out/r2g/err_c3.txt
$B --config c3 --storage 2bit > gpurun_out/r2g/bench_c3_2bit.json 2> gpurun_out/r2g/err_c3b.txt
for f in gpurun_out/r2g/bench_*.json; do echo $f; python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(' ', round(d['ms_per_step'],3),'ms kern', round(d['roofline']['kernel_ms'],3), round(d['value']/1e6,2),'M/s frac',round(d['roofline']['frac'],3), 'e2e', d.get('e2e') and round(d['e2e']['value']/1e6,2), d['config'].get('geometry'))
except Exception as ex: print('  FAILED', ex, open(sys.argv[1]).read()[-300:])
PY
done
