set -x
O=gpurun_out/r02_call4; mkdir -p $O
timeout 300 python bench.py --steps 1000 --warmup 50 --no-cpu-baseline --no-e2e > $O/bench.json 2> $O/bench.err
python -c "
import json
d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1])
print('fused', round(d['value']/1e9,3), round(d['ms_per_step']*1e3,2), d['larger_calls']['value']/1e9)
t=d['phase_trace']
for k in t['max_us']: print('  %-18s max %7.2f mean %7.2f'%(k,t['max_us'][k],t['mean_us'][k]))" || tail -5 $O/bench.err
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_batch -c 2 -o $O/ncu_kbatch python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --ncu-window 2 > $O/ncu.log 2>&1
tail -3 $O/ncu.log; ls -la $O
