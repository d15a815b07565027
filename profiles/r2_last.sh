# last check of round 2 on the shipped build: smoke(), full GPU suite, the driver's two bench commands
mkdir -p gpurun_out/r2z; cd $GRAFT_REPO_ROOT
timeout 120 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
timeout 280 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py > gpurun_out/r2z/bench_default.json 2> gpurun_out/r2z/err_default.txt
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2z/bench_reference.json 2> gpurun_out/r2z/err_reference.txt
for f in gpurun_out/r2z/bench_*.json; do echo $f; python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(' ', round(d['ms_per_step'],3),'ms', round(d['value']/1e6,2),'M/s frac',d.get('roofline') and round(d['roofline']['frac'],3), 'e2e', d.get('e2e') and round(d['e2e']['value']/1e6,2), 'launches', d.get('gpu_launches'))
except Exception as ex: print('  FAILED', ex, open(sys.argv[1]).read()[-300:])
PY
done
