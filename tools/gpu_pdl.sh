#!/bin/bash
# A/B of programmatic dependent launch for the lean step (dev tool). usage: gpu_pdl.sh TAG
TAG=${1:-pdl}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_fastpath.py tests/test_gpu_parity.py tests/test_gpu_slabs.py -m gpu -q -x > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest.log
for o in 1 0 1 0; do
  for w in c4shard c2; do
    T2D_PDL=$o timeout 300 python bench.py --workload $w --no-cpu-baseline > $OUT/bench_${w}_p$o.json 2> $OUT/bench_${w}_p$o.err
    echo "pdl=$o $w: $(python -c "import json; d=json.load(open('$OUT/bench_${w}_p$o.json')); print(d['ms_per_step'], d['kernel_ms'], d.get('fault'))")"
  done
done
