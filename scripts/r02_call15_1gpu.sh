set -x
O=gpurun_out/r02_call15; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipeline" > $O/pytest_parity.log 2>&1; echo "pytest rc=$?" >> $O/pytest_parity.log
tail -3 $O/pytest_parity.log
timeout 240 python -m pytest tests/test_gpu_p2p.py tests/test_gpu_global.py -m gpu -x -q -k "pipeline" > $O/pytest_ring.log 2>&1; echo "pytest rc=$?" >> $O/pytest_ring.log
tail -3 $O/pytest_ring.log
timeout 300 python bench.py --steps 2000 --warmup 50 --variants --no-cpu-baseline --no-traffic > $O/bench.json 2> $O/bench.err
python -c "
import json
d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1])
print('pipeline', round(d['value']/1e9,3), round(d['ms_per_step']*1e3,2), 'e2e', d['e2e']['value']/1e9)
print(d['roofline']['kernel_ms']); print(d['variants']); print(json.dumps(d['phase_trace']['isolated_batch']))" || tail -5 $O/bench.err
