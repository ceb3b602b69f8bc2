#!/bin/bash
# round-2 evidence pass on one GPU: parity suite, both bench arms, ncu launch list of the bench command, --set full
# captures of the ensemble kernel (4 M trajectories), the compaction and the statistics reduction, the C3/C4/C5 kernels
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r2_bench_ref.log 2>&1; echo "rc=$?" >> gpurun_out/r2_bench_ref.log
timeout 900 python bench.py > gpurun_out/r2_bench_1e7.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2_bench_1e7.log
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --no-policy0 > gpurun_out/r2_bench_for_ncu_plain.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --no-policy0 > gpurun_out/r2_bench_for_ncu.log 2>&1
echo "ncu-launches rc=$?"
export PDMP_TSTOP_POLICY=1
python scripts/profile_run.py tcp 4194304 > gpurun_out/r2_profile_plain.log 2>&1
for k in ensemble_kernel compact_kernel stats_reduce_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/r2_prof_$k -f python scripts/profile_run.py tcp 4194304 > gpurun_out/r2_ncu_$k.log 2>&1
  echo "ncu $k rc=$?"
done
unset PDMP_TSTOP_POLICY
for spec in "morris_lecar 1048576" "tcp2d 4194304" "sir_rejection 2097152"; do
  set -- $spec
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:ensemble_kernel -s 1 -c 1 -o gpurun_out/r2_prof_$1 -f python scripts/profile_run.py $1 $2 > gpurun_out/r2_ncu_$1.log 2>&1
  echo "ncu $1 rc=$?"
done
for f in r2_pytest_gpu r2_bench_ref r2_bench_1e7 r2_profile_plain; do echo "== $f"; tail -n 3 gpurun_out/$f.log | cut -c1-500; done
