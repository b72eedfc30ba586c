set -x
O=gpurun_out/r02_call9; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipeline" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
for v in mb3 mb4; do
  GUB_LIB=$PWD/gubernator_b200/libgub_v_$v.so timeout 300 python bench.py --steps 1500 --warmup 50 --no-cpu-baseline --no-e2e --no-traffic > $O/bench_$v.json 2> $O/bench_$v.err
  python -c "
import json
d=json.loads(open('$O/bench_$v.json').read().strip().splitlines()[-1])
print('$v', round(d['value']/1e9,3), round(d['ms_per_step']*1e3,2), d['roofline']['kernel_ms'])" || tail -3 $O/bench_$v.err
done
timeout 900 python bench.py --steps 2000 --warmup 50 --variants --no-cpu-baseline > $O/bench.json 2> $O/bench.err
python -c "
import json
d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1])
print('pipeline', round(d['value']/1e9,3), round(d['ms_per_step']*1e3,2), 'e2e', d['e2e']['value']/1e9)
print(d['roofline']); print(d['variants'])" || tail -5 $O/bench.err
