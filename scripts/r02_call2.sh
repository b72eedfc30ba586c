set -x
O=gpurun_out/r02_call2; mkdir -p $O
timeout 120 python -c "
import numpy as np, sys
sys.path[:0]=['tests','oracle']
import gubernator_b200 as g, oracle_py as O
from workloads import T0, adversarial_batch, bench_requests, zipf_ids
rng=np.random.default_rng(7)
tab=g.Table(1<<16); pool=O.Pool(workers=4, cache_size=10_000_000, now_ms=T0)
for b in range(4):
    now=T0+1000*b; pool.set_now(now)
    reqs = bench_requests(zipf_ids(rng, 8192 if b<3 else 200000, 20000, 1.1), now) if b!=2 else adversarial_batch(rng, 8192, 60, now)
    got=tab.submit(reqs, g.clock_fill(now)); want=pool.submit_hashed(reqs)
    print('batch',b,'equal',np.array_equal(got,want), 'diff', int((got!=want).sum()))
print(tab.counters())
" > $O/first.log 2>&1
cat $O/first.log
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -15 $O/pytest.log
for v in fused fused_nopdl fused_nocoop legacy; do
  case $v in fused) E="";; fused_nopdl) E="GUB_PDL=0";; fused_nocoop) E="GUB_COOP=0";; legacy) E="GUB_FUSED=0";; esac
  env $E timeout 300 python bench.py --steps 2000 --warmup 50 --no-cpu-baseline > $O/bench_$v.json 2> $O/bench_$v.err
  python -c "
import json
d=json.loads(open('$O/bench_$v.json').read().strip().splitlines()[-1])
print('$v', round(d['value']/1e9,3), round(d['ms_per_step']*1e3,2), {k:round(x*1e3,2) for k,x in d['roofline']['kernel_ms'].items()}, 'e2e', round(d['e2e']['value']/1e9,3), d['larger_calls']['value']/1e9, d['counters'])" || tail -5 $O/bench_$v.err
done
