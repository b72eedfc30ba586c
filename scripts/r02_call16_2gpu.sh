set -x
O=gpurun_out/r02_call16; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/smi.txt
timeout 200 python -m pytest tests/test_gpu_p2p.py tests/test_gpu_sharded.py -m gpu -x -q -k "pipeline or sharded" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
run() {
  name=$1; shift
  env GUB_BENCH_PROGRESS=1 GUB_BENCH_WATCHDOG=140 "$@" > $O/$name.json 2> $O/$name.err
  echo "$name rc=$?"
  python -c "
import json
d=json.loads(open('$O/$name.json').read().strip().splitlines()[-1])
print('$name', round(d['value']/1e9,3), round(d['ms_per_step']*1e3,2), 'e2e', (d.get('e2e') or {}).get('value'), d.get('ring_error'), d['roofline']['kernel_ms'], d.get('global'))" 2>/dev/null || (grep "bench rank" $O/$name.err | tail -4; tail -12 $O/$name.err)
}
TR="timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run zipf2 $TR --master-port 29711 bench.py --gpus 2 --steps 1500 --warmup 50
run global2 $TR --master-port 29712 bench.py --gpus 2 --workload global --steps 2000 --warmup 50 --no-e2e
