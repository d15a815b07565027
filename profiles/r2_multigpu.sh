# 2-GPU evidence: cross-GPU sharded parity test, the default bench (2 independent chains + checked row-sharded leg), C4 traits one per GPU
mkdir -p gpurun_out/r2mg; cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_sharded.py -x -q > gpurun_out/r2mg/pytest_sharded_2gpu.txt 2>&1; tail -3 gpurun_out/r2mg/pytest_sharded_2gpu.txt
T="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$T bench.py --gpus 2 --steps 20 --warmup 10 --no-cpu > gpurun_out/r2mg/bench_c2_n2.json 2> gpurun_out/r2mg/err_c2_n2.txt; tail -c 900 gpurun_out/r2mg/bench_c2_n2.json
$T bench.py --gpus 2 --config c4 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2mg/bench_c4_n2.json 2> gpurun_out/r2mg/err_c4_n2.txt; tail -c 400 gpurun_out/r2mg/bench_c4_n2.json
$T bench.py --gpus 2 --config c5 --sharded --steps 3 --warmup 3 --no-cpu > gpurun_out/r2mg/bench_c5_sharded_n2.json 2> gpurun_out/r2mg/err_c5s_n2.txt; tail -c 300 gpurun_out/r2mg/bench_c5_sharded_n2.json
