#!/bin/bash
# One GPU-box session: parity tests, bench, ncu launch list, ncu full capture of the step's kernels.
# Usage (here): gpurun --timeout 1500 -- 'bash tools/gpu_round.sh [tag]'
TAG=${1:-r2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt
echo "== pytest" ; timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest.log; tail -5 $OUT/pytest.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log; tail -3 $OUT/smoke.log
echo "== bench" ; timeout 600 python bench.py --steps 50 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -3 $OUT/bench.err
for W in c2 c5; do timeout 300 python bench.py --workload $W --steps 20 --warmup 3 > $OUT/bench_$W.json 2> $OUT/bench_$W.err; echo "bench $W rc=$?"; cat $OUT/bench_$W.json; done
timeout 300 python bench.py --dtype f64 --steps 20 --warmup 3 > $OUT/bench_f64.json 2> $OUT/bench_f64.err; echo "bench f64 rc=$?"; cat $OUT/bench_f64.json | cut -c1-400
echo "== bench reference" ; timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cat $OUT/bench_ref.json
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 10 --warmup 50 --no-cpu-baseline > $OUT/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_step_fast2|k_scatter_lean|k_scan_onepass' -s 150 -c 6 -f -o $OUT/prof python bench.py --steps 2 --warmup 50 --no-cpu-baseline > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i $OUT/prof.ncu-rep --page raw --csv > $OUT/raw.csv 2>/dev/null
ncu -i $OUT/prof.ncu-rep --page source --csv --print-source cuda,sass > $OUT/source.csv 2>/dev/null
ls -la $OUT
