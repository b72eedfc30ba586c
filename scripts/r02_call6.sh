set -x
O=gpurun_out/r02_call6; mkdir -p $O
timeout 300 python bench.py --steps 1000 --warmup 50 --no-cpu-baseline --no-e2e > $O/bench.json 2> $O/bench.err
python -c "
import json
d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1])
print('fused', round(d['value']/1e9,3), round(d['ms_per_step']*1e3,2), 'big', d['larger_calls']['value']/1e9)
t=d['phase_trace']
for k in t['max_us']: print('  %-18s max %7.2f mean %7.2f'%(k,t['max_us'][k],t['mean_us'][k]))
print(t['slowest_cta']); print(t['median_cta_us'])" || tail -5 $O/bench.err
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_p2p.py tests/test_gpu_global.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -15 $O/pytest.log
