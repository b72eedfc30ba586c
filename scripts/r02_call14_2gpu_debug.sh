set -x
O=gpurun_out/r02_call14; mkdir -p $O
run() {  # name, env..., then bench args
  name=$1; shift
  env GUB_BENCH_PROGRESS=1 GUB_BENCH_WATCHDOG=100 "$@" > $O/$name.json 2> $O/$name.err
  echo "$name rc=$?"
  grep "bench rank" $O/$name.err | tail -6
  python -c "
import json
d=json.loads(open('$O/$name.json').read().strip().splitlines()[-1])
print('$name', round(d['value']/1e9,3), round(d['ms_per_step']*1e3,2), 'e2e', (d.get('e2e') or {}).get('value'), d.get('ring_error'), d['roofline']['kernel_ms'], d.get('global'))" 2>/dev/null || tail -12 $O/$name.err
}
TR="timeout 130 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run s3 GUB_RING_STREAMS=3 $TR --master-port 29711 bench.py --gpus 2 --keys 4000000 --steps 400 --warmup 30
run s2 GUB_RING_STREAMS=2 $TR --master-port 29712 bench.py --gpus 2 --keys 4000000 --steps 400 --warmup 30
run s3c GUB_RING_STREAMS=3 CUDA_DEVICE_MAX_CONNECTIONS=32 $TR --master-port 29713 bench.py --gpus 2 --keys 4000000 --steps 400 --warmup 30 --no-e2e
run g2 GUB_RING_STREAMS=2 $TR --master-port 29714 bench.py --gpus 2 --workload global --keys 2000000 --steps 600 --warmup 30 --no-e2e
