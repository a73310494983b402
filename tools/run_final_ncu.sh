# final ncu evidence for the committed kernels (run AFTER the same commands exited 0 without ncu)
set -x
export LOWEST=64
for m in vcycle_rq_gs vcycle_rq_wj; do
  timeout 120 python tools/profile_sweep.py $m 4096 > gpurun_out/plain_$m.log 2>&1 || exit 1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:uni5_leg_kernel -c 2 -f -o gpurun_out/r2f_$m python tools/profile_sweep.py $m 4096 > gpurun_out/ncu_$m.log 2>&1
done
timeout 120 python tools/profile_sweep.py gsdown1 4096 > gpurun_out/plain_gsdown1.log 2>&1 || exit 1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:uni9_leg_kernel -c 1 -f -o gpurun_out/r2f_uni9_gsdown1 python tools/profile_sweep.py gsdown1 4096 > gpurun_out/ncu_gsdown1.log 2>&1
timeout 200 python bench.py --steps 4 --warmup 3 --no-side --no-cpu --e2e-steps 0 > gpurun_out/plain_step.log 2>&1 || exit 1
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2f_launches_bench_step.csv python bench.py --steps 4 --warmup 3 --no-side --no-cpu --e2e-steps 0 > gpurun_out/ncu_step.log 2>&1
ls -la gpurun_out | tail -8
