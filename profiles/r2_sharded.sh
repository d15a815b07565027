# row-sharded single chain on the blocked kernel over N GPUs (N = $1): C5 and C2 sweeps, and the default bench with its checked legs
N=$1; mkdir -p gpurun_out/r2sh$N; cd $GRAFT_REPO_ROOT
T="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513"
$T bench.py --gpus $N --config c5 --sharded --steps 20 --warmup 5 --no-cpu > gpurun_out/r2sh$N/bench_c5_sharded_blocked_n$N.json 2> gpurun_out/r2sh$N/err_c5s.txt; tail -2 gpurun_out/r2sh$N/err_c5s.txt
$T bench.py --gpus $N --config c2 --sharded --steps 20 --warmup 5 --no-cpu > gpurun_out/r2sh$N/bench_c2_sharded_blocked_n$N.json 2> gpurun_out/r2sh$N/err_c2s.txt; tail -2 gpurun_out/r2sh$N/err_c2s.txt
$T bench.py --gpus $N --config c5 --steps 10 --warmup 5 --no-cpu > gpurun_out/r2sh$N/bench_c5_n$N.json 2> gpurun_out/r2sh$N/err_c5.txt; tail -2 gpurun_out/r2sh$N/err_c5.txt
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2sh$N/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'ms_per_step',round(d['ms_per_step'],3),'kernel_ms',round(d['roofline']['kernel_ms'],3),'value M/s',round(d['value']/1e6,2), 'sharded' in d and {k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk in ('per_marker_us','max_rel_err','ranks_identical')}) for k,v in d['sharded'].items() if k in ('per_marker_us','max_rel_err','ranks_identical','blocked','error')})
    except Exception as ex: print(f,'FAILED',ex)
PY
