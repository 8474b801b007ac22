#!/bin/bash
# A/B differently tuned builds: tools/ab.sh lib1.so lib2.so ...   (prints fps and stage times of the default bench)
for lib in "$@"; do
  VELLO_B200_LIB=$PWD/$lib python bench.py --steps 30 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms']
print('$lib', round(d['value'],1), {k: round(v,3) for k,v in s.items()})"
done
