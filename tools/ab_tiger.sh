#!/bin/bash
# like tools/ab.sh, on the tiger at 1920x1080 area AA (BASELINE config C2) and at 4096^2 MSAA16
for lib in "$@"; do
  for cfg in "--size 1920 --height 1080 --aa 0" "--size 4096 --aa 2"; do
  VELLO_B200_LIB=$PWD/$lib python bench.py --scene tiger $cfg --steps 50 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms']
print('$lib', d['config']['workload'], round(d['value'],1), round(d['e2e']['value'],1), {k: round(v,3) for k,v in s.items()})"
  done
done
