set -x
O=gpurun_out/r02_call13; mkdir -p $O
timeout 420 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipeline" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -4 $O/pytest.log
timeout 400 python bench.py --steps 2000 --warmup 50 --variants --no-cpu-baseline --no-traffic > $O/bench.json 2> $O/bench.err
python -c "
import json
d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1])
print('pipeline', round(d['value']/1e9,3), round(d['ms_per_step']*1e3,2), 'e2e', d['e2e']['value']/1e9)
print(d['roofline']['kernel_ms']); print(d['variants']); print(json.dumps(d['phase_trace']))" || tail -5 $O/bench.err
GUB_LIB=$PWD/gubernator_b200/libgub_v_fin1.so timeout 240 python bench.py --steps 1000 --warmup 50 --variants --no-cpu-baseline --no-traffic --no-e2e > $O/bench_fin1.json 2> $O/bench_fin1.err
python -c "
import json
d=json.loads(open('$O/bench_fin1.json').read().strip().splitlines()[-1])
print('fin1', round(d['value']/1e9,3), round(d['ms_per_step']*1e3,2)); print(d['variants'])" || tail -5 $O/bench_fin1.err
GUB_BENCH_WATCHDOG=200 timeout 240 python bench.py --workload global --steps 1500 --warmup 50 --no-e2e > $O/bench_global1.json 2> $O/bench_global1.err
python -c "
import json
d=json.loads(open('$O/bench_global1.json').read().strip().splitlines()[-1])
print('global N=1', round(d['value']/1e9,3), round(d['ms_per_step']*1e3,2), d['global'], d.get('ring_error'))" || tail -25 $O/bench_global1.err
