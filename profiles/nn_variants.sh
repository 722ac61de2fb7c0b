#!/bin/bash
# linkage-kernel variants at the bench workload (diagnostic; not part of the bench contract)
run() { env "$@" timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --part1-only 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(' '.join(sys.argv[1:]), round(d['stage_ms_part1']['linkage'],2), d['linkage_stats']['searches'], d['linkage_stats']['cached_searches'])" "$@"; }
for v in "$@"; do run $(echo $v | tr ',' ' '); done
