# 2-GPU evidence of the row-sharded BLOCKED kernel: cross-GPU parity tests, the default bench with both checked sharded legs, sharded C5 / C2 sweeps
mkdir -p gpurun_out/r2mg2; cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_sharded.py -x -q > gpurun_out/r2mg2/pytest_sharded_2gpu.txt 2>&1; tail -3 gpurun_out/r2mg2/pytest_sharded_2gpu.txt
T="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
$T bench.py --gpus 2 --steps 20 --warmup 10 --no-cpu > gpurun_out/r2mg2/bench_c2_n2.json 2> gpurun_out/r2mg2/err_c2_n2.txt; tail -c 1500 gpurun_out/r2mg2/bench_c2_n2.json
$T bench.py --gpus 2 --config c5 --sharded --steps 10 --warmup 5 --no-cpu > gpurun_out/r2mg2/bench_c5_sharded_blocked_n2.json 2> gpurun_out/r2mg2/err_c5s.txt; tail -c 400 gpurun_out/r2mg2/bench_c5_sharded_blocked_n2.json; tail -2 gpurun_out/r2mg2/err_c5s.txt
$T bench.py --gpus 2 --config c2 --sharded --steps 10 --warmup 5 --no-cpu > gpurun_out/r2mg2/bench_c2_sharded_blocked_n2.json 2> gpurun_out/r2mg2/err_c2s.txt; tail -c 400 gpurun_out/r2mg2/bench_c2_sharded_blocked_n2.json; tail -2 gpurun_out/r2mg2/err_c2s.txt
