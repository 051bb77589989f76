#!/bin/bash
# GPU call 12: confirm the 2 GiB burst default on both data patterns at full scale, twice each (run-to-run spread).
set -u
OUT=gpurun_out/call12
mkdir -p $OUT
: > $OUT/summary.txt
for tag in pos1 ones1 pos2; do
  PAT=pos; case $tag in ones*) PAT=ones;; esac
  echo "== bench ours full scale, pattern $PAT ($tag)" | tee -a $OUT/summary.txt
  timeout 1200 python bench.py --pattern $PAT --keep $OUT/$tag > $OUT/$tag.json 2> $OUT/$tag.err; echo "rc=$?" | tee -a $OUT/summary.txt
  python - <<PY | tee -a $OUT/summary.txt
import json
r=json.loads(open('$OUT/$tag.json').read().strip().splitlines()[-1])
print('$tag:', 'verified', r.get('verified'), 'stall', round(r.get('stall_ms_per_handoff',0)), 'ms iter/s', round(r.get('iter_per_s',0),2), 'e2e', round(r['e2e']['value'],1), 'device', r.get('device',{}).get('evict_GBps'), r.get('device',{}).get('fetch_GBps'), 'wait', r.get('device',{}).get('wait_ms_mean'), 'gaps', [round(x) for x in r['first_iter_gap_ms']])
PY
done
find $OUT -name "client*.jsonl" -size +2M -exec truncate -s 2M {} \;
