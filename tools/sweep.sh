#!/bin/bash
# usage: tools/sweep.sh "ENV=.. ENV2=.." ...   -> prints sectors/s for each environment setting
for cfg in "$@"; do
  r=$(env $cfg python bench.py --steps 10 --warmup 3 --cpu-sample 2 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['chain_hbm_frac'],4), round(d['e2e']['value']))")
  echo "$cfg => $r"
done
