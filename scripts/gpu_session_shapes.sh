#!/bin/bash
# Which register-resident generation / seam protocol for which species count: short plans over shapes other than C5
# (scripts/multi_shapes.py: kernel time from the library's CUDA events).  Every line under its own timeout.
mkdir -p gpurun_out
out=gpurun_out/r2_reg3_shapes.jsonl
: > $out
shapes="8,128,8192,512 8,64,8192,512 7,200,4096,512 6,256,4096,512 5,300,4096,512 5,64,8192,512 4,200,8192,512 4,512,4096,512 3,100,16384,1000 2,100,16384,2000 2,512,8192,512 1,100,16384,2000"
for sh in $shapes; do
  n=${sh%%,*}
  variants="CHROM_REG2=2 CHROM_REG2=3,CHROM_REG3_SYNC=0 CHROM_REG2=3,CHROM_REG3_SYNC=1 CHROM_REG2=3,CHROM_REG3_SYNC=3"
  [ "$n" -le 4 ] && variants="CHROM_REG2=0 $variants"
  for v in $variants; do
    timeout 60 python scripts/multi_shapes.py "$sh,$v" >> $out 2>> gpurun_out/r2_reg3_shapes.err || echo "{\"failed\": \"$sh,$v\"}" >> $out
  done
done
python - <<'PY'
import json
for ln in open("gpurun_out/r2_reg3_shapes.jsonl"):
    d = json.loads(ln)
    if "failed" in d:
        print("FAILED", d["failed"]); continue
    print(d["n"], d["nz"], d["cols"], d["env"], d["kernel"].split(">")[0], round(d["kernel_ms"], 3), "%.3e" % d["cell_stage_per_s"], d["bad_status"])
PY
timeout 200 python -m pytest tests/test_gpu_multi_reg3.py -q --timeout 100 --timeout-method thread > gpurun_out/reg3_tests_default.log 2>&1; echo "pytest rc=$?" >> gpurun_out/reg3_tests_default.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/reg3_tests_default.log; tail -4 gpurun_out/smoke.log
