#!/bin/bash
# (Provenance of profiles/r2_reg3_sessions.txt.  Seam protocols 1, 2, 5 and 6 named here were measured and then removed from
# the build: CHROM_REG3_SYNC now accepts 0, 3 and 4; see csrc/multi_reg3.cuh and profiles/r2_c5_experiments.txt.)
# Sixth reg3 session: early barrier tests + late last post (CHROM_REG3_SYNC=6) against the default (4) on C5.
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_multi_reg3.py -q --timeout 60 --timeout-method thread > gpurun_out/reg3_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/reg3_tests.log
: > gpurun_out/reg3_c5.jsonl
for s in 6 4; do
  echo "# CHROM_REG2=3 CHROM_REG3_SYNC=$s" >> gpurun_out/reg3_c5.jsonl
  CHROM_REG2=3 CHROM_REG3_SYNC=$s timeout 100 python bench.py --workload c5 --steps 3 --warmup 1 --no-cpu --no-e2e >> gpurun_out/reg3_c5.jsonl 2>> gpurun_out/reg3_c5.err
  echo "# rc=$?" >> gpurun_out/reg3_c5.jsonl
done
tail -3 gpurun_out/reg3_tests.log; cut -c1-260 gpurun_out/reg3_c5.jsonl
