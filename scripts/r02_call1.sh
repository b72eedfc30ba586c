set -x
mkdir -p gpurun_out/r02_call1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02_call1/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_call1/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_call1/pytest.log
for v in base onepass cs both; do
  if [ $v = base ]; then L=""; else L=$PWD/gubernator_b200/libgub_v_$v.so; fi
  GUB_LIB=$L timeout 300 python bench.py --steps 2000 --warmup 50 --no-cpu-baseline --no-e2e > gpurun_out/r02_call1/bench_$v.json 2> gpurun_out/r02_call1/bench_$v.err
done
tail -3 gpurun_out/r02_call1/pytest.log
for v in base onepass cs both; do python -c "
import json,sys
d=json.loads(open('gpurun_out/r02_call1/bench_$v.json').read().strip().splitlines()[-1])
print('$v', d['value']/1e9, d['ms_per_step'], d['roofline']['kernel_ms'])"; done
