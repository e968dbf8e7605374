#!/bin/bash
# usage: tools/sweep.sh "ENVVAR=a ENVVAR2=b" ... ; runs a short bench for every env combination
for combo in "$@"; do
  out=$(env $combo python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu-baseline 2>&1 | tail -1)
  echo "$combo :: $(echo "$out" | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("%.1f Mrays/s  extend %.1f ms shade %.1f other %.1f" % (d["value"], d["kernel_ms"]["extend"], d["kernel_ms"]["shade"], d["kernel_ms"]["other"]))' 2>&1 | tail -1)"
done
