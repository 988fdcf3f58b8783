#!/bin/bash
# A/B of the heavy-rows-first order of k_step_fast2 (dev tool). usage: gpu_order.sh TAG
TAG=${1:-order}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_fastpath.py tests/test_gpu_parity.py tests/test_gpu_slabs.py -m gpu -q -x > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
for o in 1 0 1 0; do
  T2D_F2_ORDER=$o timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > $OUT/bench_o$o.json 2> $OUT/bench_o$o.err
  echo "order=$o $(python -c "import json; d=json.load(open('$OUT/bench_o$o.json')); print(d['ms_per_step'], d['kernel_ms'], d['e2e']['value'], d.get('fault_mask'))")"
done
cp 2dtissue_b200/lib2dtissue_b200.so /tmp/main.so
timeout 300 python tools/diag_timeline.py variants/lib_timeline.so 2>&1 | grep -A12 "^rep 2"
T2D_F2_ORDER=0 timeout 300 python tools/diag_timeline.py variants/lib_timeline.so 2>&1 | grep -A5 "^rep 2"
cp /tmp/main.so 2dtissue_b200/lib2dtissue_b200.so
