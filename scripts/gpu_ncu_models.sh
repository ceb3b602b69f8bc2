#!/bin/bash
# --set full captures of the C3 / C5 / C4 kernels (profile_run.py workloads)
mkdir -p gpurun_out
for spec in "morris_lecar 1048576" "tcp2d 4194304" "sir_rejection 2097152"; do
  set -- $spec
  python scripts/profile_run.py $1 $2 > gpurun_out/r2m_$1_plain.log 2>&1 && \
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:ensemble_kernel -s 1 -c 1 -o gpurun_out/r2m_$1 -f python scripts/profile_run.py $1 $2 > gpurun_out/r2m_$1_ncu.log 2>&1
  echo "$1 rc=$?"; tail -n 1 gpurun_out/r2m_$1_plain.log | cut -c1-330
done
