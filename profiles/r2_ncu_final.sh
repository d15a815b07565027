# ncu --set full of the sweep kernel at the default bench command (C2 int8), shipped build of round 2; the plain command runs first
mkdir -p gpurun_out/r2ncuf; cd $GRAFT_REPO_ROOT
CMD="python bench.py --no-cpu --no-e2e --steps 2 --warmup 3 --long-seconds 0"
timeout 200 $CMD > gpurun_out/r2ncuf/plain_c2.log 2>&1 && timeout 400 ncu --set full --clock-control none --import-source on -k regex:gibbs_kernel -s 4 -c 1 -o gpurun_out/r2ncuf/full_c2 $CMD > gpurun_out/r2ncuf/ncu_full_c2.log 2>&1
ls -la gpurun_out/r2ncuf/; tail -3 gpurun_out/r2ncuf/ncu_full_c2.log
