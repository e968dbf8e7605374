#!/bin/bash
# usage: tools/fsweep.sh "<bench args>" ... ; one short bench per argument string
for a in "$@"; do
  out=$(python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu-baseline $a 2>&1 | tail -1)
  echo "$a :: $(echo "$out" | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("%.1f Mrays/s  extend %.1f shade %.1f other %.1f ms" % (d["value"], d["kernel_ms"]["extend"], d["kernel_ms"]["shade"], d["kernel_ms"]["other"]))' 2>&1 | tail -1)"
done
