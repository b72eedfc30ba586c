set -x
O=gpurun_out/r02_call3; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
for v in fused sweep0; do
  case $v in fused) E="";; sweep0) E="GUB_SWEEP=0";; esac
  env $E timeout 300 python bench.py --steps 1000 --warmup 50 --no-cpu-baseline --no-e2e > $O/bench_$v.json 2> $O/bench_$v.err
  python -c "
import json
d=json.loads(open('$O/bench_$v.json').read().strip().splitlines()[-1])
print('$v', round(d['value']/1e9,3), round(d['ms_per_step']*1e3,2), d['larger_calls']['value']/1e9)
print(d['phase_trace'])" || tail -5 $O/bench_$v.err
done
