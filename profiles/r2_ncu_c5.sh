# new big-panel fall-back tests, then ncu --set full of the sweep kernel at C5 int8 (shared-tile instantiation); the plain command runs first
mkdir -p gpurun_out/r2ncu5; cd $GRAFT_REPO_ROOT
timeout 120 python -m pytest tests/test_gpu_joint.py tests/test_gpu_bayesr.py -x -q -k "native_chain" 2>&1 | tail -3
CMD="python bench.py --config c5 --no-cpu --no-e2e --steps 2 --warmup 3 --long-seconds 0"
timeout 120 $CMD > gpurun_out/r2ncu5/plain_c5.log 2>&1 && timeout 200 ncu --set full --clock-control none --import-source on -k regex:gibbs_kernel -s 4 -c 1 -o gpurun_out/r2ncu5/full_c5 $CMD > gpurun_out/r2ncu5/ncu_full_c5.log 2>&1
ls -la gpurun_out/r2ncu5/; tail -2 gpurun_out/r2ncu5/ncu_full_c5.log | cut -c1-300
