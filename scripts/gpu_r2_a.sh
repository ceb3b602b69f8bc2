#!/bin/bash
# round-2 pass A: baseline parity tests (incl. the Julia-stream goldens), PCIe peaks, ncu at 4 M trajectories
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2a_box.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
timeout 300 python scripts/pcie_probe.py > gpurun_out/r2a_pcie.log 2>&1
timeout 300 python scripts/profile_run.py tcp 4194304 > gpurun_out/r2a_profile_plain.log 2>&1 && \
PDMP_TSTOP_POLICY=1 timeout 300 python scripts/profile_run.py tcp 4194304 >> gpurun_out/r2a_profile_plain.log 2>&1 && \
PDMP_TSTOP_POLICY=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:ensemble_kernel -s 1 -c 1 -o gpurun_out/r2a_prof_tcp_4M -f python scripts/profile_run.py tcp 4194304 > gpurun_out/r2a_ncu_tcp.log 2>&1
echo "ncu rc=$?"
for f in r2a_pytest r2a_pcie r2a_profile_plain; do echo "== $f"; tail -n 5 gpurun_out/$f.log | cut -c1-700; done
