#!/bin/bash
# A/B of K1 tuning knobs on one B200: each line = environment settings, then the short bench's whole-job value,
# K1 roofline fraction, K1 launch time and the parity verdict. Usage: tools/ab_k1.sh "VAR=1 VAR2=2" "..." ...
# (an empty string = defaults). Output: gpurun_out/ab_k1.txt
mkdir -p gpurun_out
out=gpurun_out/ab_k1.txt
: > $out
for cfg in "$@"; do
  line=$(env $cfg timeout 120 python bench.py --frames ${AB_FRAMES:-2048} --chunk ${AB_CHUNK:-512} --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --no-content --sustain-seconds 0 --base-frames 8 --no-overlap 2>/dev/null | tail -1)
  echo "$line" | python -c "
import sys, json
cfg = sys.argv[1]
try:
    d = json.loads(sys.stdin.read())
    r = d['roofline']
    print('%-44s value %8.1f Gpix/s  k1 frac %.4f  k1 %.4f ms  k2 %.3f ms  parity %s  clk %s' % (cfg or '(defaults)', d['value']/1e3, r['frac'], r['avg_launch_ms'], r['k2_avg_launch_ms'], d['parity']['identical_to_oracle'], d['clocks']['sm_mhz']))
except Exception as e:
    print('%-44s FAILED %s' % (cfg, e))
" "$cfg" >> $out
done
cat $out
