#!/bin/bash
# A/B differently tuned builds: tools/ab.sh lib1.so lib2.so ...   (prints fps and stage times of the default bench;
# extra bench arguments through BENCH_ARGS, e.g. BENCH_ARGS="--scene tiger --size 1920 --height 1080 --aa 0")
for lib in "$@"; do
  VELLO_B200_LIB=$PWD/$lib python bench.py --steps 20 --no-cpu-baseline $BENCH_ARGS 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms']
print('$lib', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k: round(v,3) for k,v in s.items()})"
done
