set -x
O=gpurun_out/r02_call12; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/smi.txt
timeout 1500 python -m pytest tests/test_gpu_p2p.py tests/test_gpu_global.py tests/test_gpu_sharded.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -8 $O/pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 2000 --warmup 50 > $O/bench_2gpu.json 2> $O/bench_2gpu.err
python -c "
import json
d=json.loads(open('$O/bench_2gpu.json').read().strip().splitlines()[-1])
print('2gpu', round(d['value']/1e9,3), round(d['ms_per_step']*1e3,2), 'e2e', d['e2e']['value']/1e9, d.get('ring_error'), d['roofline']['kernel_ms'])" || tail -5 $O/bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --workload global --steps 3000 --warmup 100 --no-e2e > $O/bench_2gpu_global.json 2> $O/bench_2gpu_global.err
python -c "
import json
d=json.loads(open('$O/bench_2gpu_global.json').read().strip().splitlines()[-1])
print('2gpu global', round(d['value']/1e9,3), round(d['ms_per_step']*1e3,2), d['global'], d.get('ring_error'))" || tail -5 $O/bench_2gpu_global.err
