set -x
O=gpurun_out/r02_call11; mkdir -p $O
# source-level capture: one warm launch of each pipeline kernel (config 3), warm caches kept (--cache-control none)
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on --profile-from-start off \
  -k regex:'k_(group|rank|eval|finish)' -c 8 -o $O/pipeline_full \
  python bench.py --traffic-probe --keys 100000000 --zipf 1.1 --pool 32 > $O/ncu.log 2>&1
tail -3 $O/ncu.log
ls -la $O
