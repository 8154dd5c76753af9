#!/bin/bash
# (Provenance of profiles/r2_reg3_sessions.txt.  Seam protocols 1, 2, 5 and 6 named here were measured and then removed from
# the build: CHROM_REG3_SYNC now accepts 0, 3 and 4; see csrc/multi_reg3.cuh and profiles/r2_c5_experiments.txt.)
# One gpurun call: parity of k_multi_reg3 against k_multi_reg2, C5 timing of every seam protocol, ncu captures of the
# two mbarrier variants, then the whole GPU suite and the default bench line.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi_reg3.py -x -q > gpurun_out/reg3_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/reg3_tests.log
: > gpurun_out/reg3_c5.jsonl
for v in "2 0" "3 0" "3 1" "3 2"; do
  set -- $v
  echo "# CHROM_REG2=$1 CHROM_REG3_SYNC=$2" >> gpurun_out/reg3_c5.jsonl
  CHROM_REG2=$1 CHROM_REG3_SYNC=$2 timeout 300 python bench.py --workload c5 --steps 3 --warmup 1 --no-cpu --no-e2e >> gpurun_out/reg3_c5.jsonl 2>> gpurun_out/reg3_c5.err
done
for s in 1 2; do
  CHROM_REG2=3 CHROM_REG3_SYNC=$s timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_multi_reg3 -c 1 \
    -f -o gpurun_out/ncu_reg3_sync$s python bench.py --workload c5 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/ncu_reg3_sync$s.log 2>&1
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -3 gpurun_out/reg3_tests.log; cat gpurun_out/reg3_c5.jsonl | cut -c1-400; tail -3 gpurun_out/gputests.log
