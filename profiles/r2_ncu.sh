# ncu evidence of round 2 (one gpurun call): launch list of the default bench command, full captures of the sweep kernel at C2 (int8, 2-bit), C5, C3
mkdir -p gpurun_out/r2ncu; cd $GRAFT_REPO_ROOT
CMD="python bench.py --no-cpu --no-e2e --steps 2 --warmup 3 --long-seconds 0"
timeout 200 $CMD > gpurun_out/r2ncu/plain_c2.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2ncu/launches_c2.csv $CMD > gpurun_out/r2ncu/ncu_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gibbs_kernel -s 4 -c 1 -o gpurun_out/r2ncu/full_c2 $CMD > gpurun_out/r2ncu/ncu_full_c2.log 2>&1
timeout 200 $CMD --storage 2bit > gpurun_out/r2ncu/plain_c2_2bit.log 2>&1 && timeout 300 ncu --set full --clock-control none -k regex:gibbs_kernel -s 4 -c 1 -o gpurun_out/r2ncu/full_c2_2bit $CMD --storage 2bit > gpurun_out/r2ncu/ncu_full_c2_2bit.log 2>&1
timeout 200 $CMD --config c5 > gpurun_out/r2ncu/plain_c5.log 2>&1 && timeout 300 ncu --set full --clock-control none -k regex:gibbs_kernel -s 4 -c 1 -o gpurun_out/r2ncu/full_c5 $CMD --config c5 > gpurun_out/r2ncu/ncu_full_c5.log 2>&1
timeout 300 $CMD --config c3 > gpurun_out/r2ncu/plain_c3.log 2>&1 && timeout 500 ncu --set full --clock-control none -k regex:gibbs_kernel -s 4 -c 1 -o gpurun_out/r2ncu/full_c3 $CMD --config c3 > gpurun_out/r2ncu/ncu_full_c3.log 2>&1
timeout 300 $CMD --config c3 --storage 2bit > gpurun_out/r2ncu/plain_c3_2bit.log 2>&1 && timeout 500 ncu --set full --clock-control none -k regex:gibbs_kernel -s 4 -c 1 -o gpurun_out/r2ncu/full_c3_2bit $CMD --config c3 --storage 2bit > gpurun_out/r2ncu/ncu_full_c3_2bit.log 2>&1
ls -la gpurun_out/r2ncu/
