#!/bin/bash
# usage: profiles/tune.sh OUT "ENV1=a ENV2=b" "ENV1=c" ...   -> one bench line per variant (kernel ms, ms_per_step)
out=$1; shift
: > $out
for v in "$@"; do
  echo "== $v" >> $out
  env $v python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-solver 2>>$out.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'ms_per_step':round(d['ms_per_step'],3),'rhs_ms':round(d['rhs_only']['ms'],3),'les_ms':round(d['les_only']['ms'],3),'kernel_ms':{k:round(v['ms'],3) for k,v in d['kernels'].items() if isinstance(v,dict)}}))" >> $out
done
cat $out
