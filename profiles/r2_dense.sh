# dense-set ring geometry: the new test, the full GPU suite, the dense bench lines
mkdir -p gpurun_out/r2d; cd $GRAFT_REPO_ROOT
timeout 60 python -m pytest tests/test_gpu_parity.py -x -q -k "dense_sets" 2>&1 | tail -3 || exit 1
timeout 280 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
B="timeout 100 python bench.py --no-cpu --steps 10 --warmup 5"
$B --config c1 --steps 20 --warmup 10 > gpurun_out/r2d/bench_c1.json 2> gpurun_out/r2d/err_c1.txt
$B --config c4rr --no-e2e > gpurun_out/r2d/bench_c4rr.json 2> gpurun_out/r2d/err_c4rr.txt
for f in gpurun_out/r2d/bench_*.json; do echo $f; python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(' ', round(d['ms_per_step'],3),'ms kern', round(d['roofline']['kernel_ms'],3), round(d['value']/1e6,2),'M/s', 'e2e', d.get('e2e') and round(d['e2e']['value']/1e6,2), d['config'].get('geometry'))
except Exception as ex: print('  FAILED', ex, open(sys.argv[1]).read()[-300:])
PY
done
