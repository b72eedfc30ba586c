set -x
O=gpurun_out/r02_call8; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_p2p.py tests/test_gpu_global.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -12 $O/pytest.log
timeout 600 python bench.py --steps 2000 --warmup 50 --variants > $O/bench.json 2> $O/bench.err
python -c "
import json
d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1])
print('pipeline', round(d['value']/1e9,3), round(d['ms_per_step']*1e3,2), 'big', d['larger_calls'], 'e2e', d['e2e']['value']/1e9, d['e2e']['prehashed_compact']['value']/1e9)
print(d['roofline']); print(d['variants']); print(d['cpu_baseline'])" || tail -5 $O/bench.err
GUB_PATH=fused timeout 300 python bench.py --steps 1000 --warmup 50 --no-cpu-baseline --no-traffic > $O/bench_fused.json 2> $O/bench_fused.err
python -c "
import json
d=json.loads(open('$O/bench_fused.json').read().strip().splitlines()[-1])
print('fused', round(d['value']/1e9,3), round(d['ms_per_step']*1e3,2), 'e2e', d['e2e']['value']/1e9)" || tail -5 $O/bench_fused.err
