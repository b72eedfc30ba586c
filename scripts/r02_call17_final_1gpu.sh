set -x
O=gpurun_out/r02_call17; mkdir -p $O
timeout 480 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
timeout 330 python bench.py > $O/bench.json 2> $O/bench.err
python -c "
import json
d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1])
print('final', round(d['value']/1e9,3), round(d['ms_per_step']*1e3,2), 'e2e', d['e2e']['value']/1e9, 'cpu', d['cpu_baseline']['value']/1e6 if d.get('cpu_baseline') else None)
print(d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['fractions'])" || tail -5 $O/bench.err
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:'k_(group|rank|eval|finish)' --csv --log-file $O/ncu_launches.csv python bench.py --traffic-probe --keys 100000000 --zipf 1.1 --pool 32 > $O/ncu.log 2>&1
tail -2 $O/ncu.log
