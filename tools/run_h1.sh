#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python bench.py --frames 1024 --width 1920 --height 1080 --steps 5 --warmup 3 --no-cpu-baseline --no-content --sustain-seconds 0 > $O/r02_cfg2_1080p.json 2> $O/h1_err.txt
timeout 600 python bench.py --frames 512 --gridn 14 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-content --sustain-seconds 0 > $O/r02_cfg4_n14_L0.json 2>> $O/h1_err.txt
timeout 600 python bench.py --frames 1024 --width 1920 --height 1080 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-content --sustain-seconds 0 --no-overlap > $O/r02_cfg2_1080p_noov.json 2>> $O/h1_err.txt
timeout 600 python tools/bench_mixed.py > $O/r02_cfg5_mixed_batch.json 2>> $O/h1_err.txt
python -c "
import json
for f in ('r02_cfg2_1080p.json','r02_cfg4_n14_L0.json','r02_cfg2_1080p_noov.json'):
    d=json.load(open('$O/'+f)); r=d['roofline']; print(f, round(d['value']/1e3), 'Gpix/s frac', round(r['frac'],4), 'step_frac', round(r['step_frac'],4), 'share', round(r['k1_share_of_step'],3), d['parity']['identical_to_oracle'])
"
cut -c1-180 $O/r02_cfg5_mixed_batch.json; tail -2 $O/h1_err.txt
