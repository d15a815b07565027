mkdir -p gpurun_out/r2shn2; cd $GRAFT_REPO_ROOT
T="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515"
$T bench.py --gpus 2 --config c5 --sharded --steps 20 --warmup 5 --no-cpu > gpurun_out/r2shn2/bench_c5_sharded_blocked_n2.json 2> gpurun_out/r2shn2/err1.txt; tail -2 gpurun_out/r2shn2/err1.txt
$T bench.py --gpus 2 --config c5 --sharded --steps 20 --warmup 5 --no-cpu --cfg-opt 258 > gpurun_out/r2shn2/bench_c5_sharded_blocked_hier_n2.json 2> gpurun_out/r2shn2/err2.txt; tail -2 gpurun_out/r2shn2/err2.txt
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2shn2/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'ms_per_step',round(d['ms_per_step'],3),'kernel_ms',round(d['roofline']['kernel_ms'],3),'value M/s',round(d['value']/1e6,2), d['config']['geometry'])
    except Exception as ex: print(f,'FAILED',ex)
PY
