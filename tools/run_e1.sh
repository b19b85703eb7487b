#!/bin/bash
# content table with the cascade (default) and with the tiled kernel (kernel_variant 2)
O=gpurun_out; mkdir -p $O
for v in 0 2; do
  timeout 600 python bench.py --frames 256 --chunk 128 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --sustain-seconds 0 --base-frames 8 --kernel-variant $v > $O/e1_v$v.json 2>> $O/e1_err.txt
  python -c "
import json,sys
d=json.load(open('$O/e1_v$v.json'))
print('variant $v value', d['value'], 'frac', d['roofline']['frac'])
for c in d['roofline']['by_content']: print('   %-70s k1 %.3f ms frac %.4f cands %d ok %s' % (c['content'][:70], c['k1_ms'], c['frac'], c['candidates_per_frame'], c['identical_to_oracle']))
"
done
tail -3 $O/e1_err.txt
