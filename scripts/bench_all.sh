#!/bin/bash
# short bench of the four named codes (frames/s at 30 iterations | operating point | e2e), for A/B comparisons
for c in ${CODES:-wifi a5 c79 a24}; do
  python bench.py --code $c --steps ${STEPS:-3} --no-cpu --frames ${FRAMES:-65536} "$@" 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(d['config']['code'], 'fps30', round(d['frames_per_s']), 'op', round(d['operating_point']['frames_per_s']), 'e2e', round(d['e2e']['value'], 3), 'val', round(d['value'], 3))"
done
