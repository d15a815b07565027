# row-sharded single chain on the blocked kernel over N GPUs with and without the rank-local pre-reduction: C5 sweeps
N=$1; mkdir -p gpurun_out/r2shh$N; cd $GRAFT_REPO_ROOT
T="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514"
$T bench.py --gpus $N --config c5 --sharded --steps 20 --warmup 5 --no-cpu > gpurun_out/r2shh$N/bench_c5_sharded_blocked_hier_n$N.json 2> gpurun_out/r2shh$N/err_c5s.txt; tail -2 gpurun_out/r2shh$N/err_c5s.txt
$T bench.py --gpus $N --config c2 --sharded --steps 20 --warmup 5 --no-cpu > gpurun_out/r2shh$N/bench_c2_sharded_blocked_hier_n$N.json 2> gpurun_out/r2shh$N/err_c2s.txt; tail -2 gpurun_out/r2shh$N/err_c2s.txt
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2shh$N/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'ms_per_step',round(d['ms_per_step'],3),'kernel_ms',round(d['roofline']['kernel_ms'],3),'value M/s',round(d['value']/1e6,2), d['config']['geometry'])
    except Exception as ex: print(f,'FAILED',ex)
PY
