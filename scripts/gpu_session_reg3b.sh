#!/bin/bash
# (Provenance of profiles/r2_reg3_sessions.txt.  Seam protocols 1, 2, 5 and 6 named here were measured and then removed from
# the build: CHROM_REG3_SYNC now accepts 0, 3 and 4; see csrc/multi_reg3.cuh and profiles/r2_c5_experiments.txt.)
# Second reg3 session: every step under its own timeout.  C5 timing of the RK4 variants, then probes that locate the
# Euler stall seen in the first session (which kernel generation / seam protocol / kind of column).
mkdir -p gpurun_out
: > gpurun_out/reg3_c5.jsonl
for v in "2 0" "3 0" "3 1" "3 2"; do
  set -- $v
  echo "# CHROM_REG2=$1 CHROM_REG3_SYNC=$2" >> gpurun_out/reg3_c5.jsonl
  CHROM_REG2=$1 CHROM_REG3_SYNC=$2 timeout 200 python bench.py --workload c5 --steps 3 --warmup 1 --no-cpu --no-e2e >> gpurun_out/reg3_c5.jsonl 2>> gpurun_out/reg3_c5.err
  echo "# rc=$?" >> gpurun_out/reg3_c5.jsonl
done
: > gpurun_out/reg3_probe.log
probe() { timeout 40 python scripts/reg3_probe.py "$@" >> gpurun_out/reg3_probe.log 2>&1; rc=$?; [ $rc -ne 0 ] && echo "RC=$rc $*" >> gpurun_out/reg3_probe.log; return $rc; }
for v in "2 0" "3 0" "3 1" "3 2"; do
  set -- $v
  if probe 8 512 euler $1 $2 mild; then
    for kind in strong late unstable all; do probe 8 512 euler $1 $2 $kind; done
  else
    probe 8 64 euler $1 $2 mild 20
    probe 3 64 euler $1 $2 mild 20
  fi
done
probe 8 512 rk4 3 1 all
timeout 300 python -m pytest tests/test_gpu_multi_reg3.py -x -q -k "rk4 or late or many" > gpurun_out/reg3_tests_rk4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/reg3_tests_rk4.log
cat gpurun_out/reg3_c5.jsonl | cut -c1-300; cat gpurun_out/reg3_probe.log; tail -3 gpurun_out/reg3_tests_rk4.log
