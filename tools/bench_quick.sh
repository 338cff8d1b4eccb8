#!/bin/bash
# headline workload only (value, roofline.frac), for A/B runs of K-B switches: tools/bench_quick.sh [label]
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-dpp --no-configs --no-incremental-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', round(d['value']), round(d['roofline']['frac'],4), d['roofline']['frac_by_n'])"
