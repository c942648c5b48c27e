#!/bin/bash
# usage: quick_bench.sh <label> [env assignments...]   - short bench line summary
label=$1; shift
env "$@" timeout 900 python bench.py --steps 2 --warmup 2 --skip-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('$label', 'value=%.2f e2e=%.2f' % (d['value'], d['e2e']['value']), {k: round(v,2) for k,v in d['stage_ms'].items()})"
