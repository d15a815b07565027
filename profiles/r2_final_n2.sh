# final 2-GPU check of round 2 on the shipped build: cross-GPU parity tests, the driver's 2-GPU bench of both arms, the row-sharded C5 / C2 chains
mkdir -p gpurun_out/r2f2; cd $GRAFT_REPO_ROOT
timeout 400 python -m pytest tests/test_gpu_sharded.py -q -x 2>&1 | tail -3 > gpurun_out/r2f2/pytest_sharded_2gpu.txt; cat gpurun_out/r2f2/pytest_sharded_2gpu.txt
T="timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515"
$T bench.py --gpus 2 --steps 10 --warmup 5 > gpurun_out/r2f2/bench_c2_n2.json 2> gpurun_out/r2f2/err0.txt; tail -2 gpurun_out/r2f2/err0.txt
$T bench.py --gpus 2 --config c5 --sharded --steps 20 --warmup 5 --no-cpu > gpurun_out/r2f2/bench_c5_sharded_blocked_hier_n2.json 2> gpurun_out/r2f2/err1.txt; tail -2 gpurun_out/r2f2/err1.txt
$T bench.py --gpus 2 --config c2 --sharded --steps 20 --warmup 5 --no-cpu > gpurun_out/r2f2/bench_c2_sharded_blocked_hier_n2.json 2> gpurun_out/r2f2/err2.txt; tail -2 gpurun_out/r2f2/err2.txt
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2f2/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'ms_per_step',round(d['ms_per_step'],3),'kernel_ms',round(d['roofline']['kernel_ms'],3),'value M/s',round(d['value']/1e6,2), d['config']['geometry'], d['config'].get('kernel_variant'), d.get('sharded_leg') or d['config'].get('sharded_leg'))
    except Exception as ex: print(f,'FAILED',ex)
PY
