#!/bin/bash
# run the default bench R times with K timed steps and report the fault mask of each run (fp32 sums depend on arrival
# order, so runs differ in the last bits: a physically marginal workload shows up as an occasional fault)
R=${1:-5}; K=${2:-100}
for r in $(seq 1 $R); do
  python bench.py --steps $K --warmup 5 --no-cpu-baseline 2> /tmp/err.txt | cut -c1-160; grep -c "fault mask" /tmp/err.txt
done
