#!/bin/bash
# dev: e2e (host-buffer) throughput of the default bench for a few chunk / thread settings of Engine::step_host32
for cfg in "4 16" "8 16" "2 16" "4 8" "4 32" "8 32"; do set -- $cfg
  T2D_HOST32_K=$1 T2D_HOST32_T=$2 python bench.py 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('K=$1 T=$2', d['e2e']['value'])"
done
nproc; lscpu | grep -E "Model name|Socket|NUMA node\(s\)|^CPU\(s\)"
