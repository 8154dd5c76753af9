#!/bin/bash
# (Provenance of profiles/r2_reg3_sessions.txt.  Seam protocols 1, 2, 5 and 6 named here were measured and then removed from
# the build: CHROM_REG3_SYNC now accepts 0, 3 and 4; see csrc/multi_reg3.cuh and profiles/r2_c5_experiments.txt.)
# Third reg3 session (mbarriers initialised once, phase bookkeeping across runs): probes of the shapes that stalled,
# C5 timing of the mbarrier variants, the whole GPU suite, the default bench line, one full ncu capture.
mkdir -p gpurun_out
: > gpurun_out/reg3_probe.log
probe() { timeout 40 python scripts/reg3_probe.py "$@" >> gpurun_out/reg3_probe.log 2>&1; rc=$?; [ $rc -ne 0 ] && echo "RC=$rc $*" >> gpurun_out/reg3_probe.log; return $rc; }
ok1=1; ok2=1
probe 8 512 euler 3 1 strong || ok1=0
probe 8 512 euler 3 1 all || ok1=0
probe 8 512 rk4 3 1 all || ok1=0
probe 8 64 rk4 3 1 all 300 || ok1=0
probe 8 512 euler 3 2 all || ok2=0
probe 8 512 rk4 3 2 all || ok2=0
probe 8 64 rk4 3 2 all 300 || ok2=0
: > gpurun_out/reg3_c5.jsonl
for s in 1 2; do
  eval ok=\$ok$s
  echo "# CHROM_REG2=3 CHROM_REG3_SYNC=$s probes_ok=$ok" >> gpurun_out/reg3_c5.jsonl
  if [ "$ok" = 1 ]; then
    CHROM_REG2=3 CHROM_REG3_SYNC=$s timeout 150 python bench.py --workload c5 --steps 3 --warmup 1 --no-cpu --no-e2e >> gpurun_out/reg3_c5.jsonl 2>> gpurun_out/reg3_c5.err
    echo "# rc=$?" >> gpurun_out/reg3_c5.jsonl
  fi
done
timeout 700 python -m pytest tests -m gpu -q --timeout 300 --timeout-method thread --ignore=tests/test_gpu_multi_reg3.py > gpurun_out/gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputests.log
timeout 400 python -m pytest tests/test_gpu_multi_reg3.py -q --timeout 100 --timeout-method thread > gpurun_out/reg3_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/reg3_tests.log
timeout 500 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?" >> gpurun_out/bench_default.err
CHROM_REG2=3 CHROM_REG3_SYNC=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_multi_reg3 -c 1 \
  -f -o gpurun_out/ncu_reg3_sync0 python bench.py --workload c5 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/ncu_reg3_sync0.log 2>&1
cat gpurun_out/reg3_probe.log; cut -c1-260 gpurun_out/reg3_c5.jsonl; tail -4 gpurun_out/gputests.log; tail -4 gpurun_out/reg3_tests.log; tail -2 gpurun_out/bench_default.err; tail -2 gpurun_out/ncu_reg3_sync0.log
