#!/bin/bash
# launch-order experiment: fs time (CUDA events, plain run) and DRAM traffic (ncu metrics pass) per setting
mkdir -p gpurun_out
CMD="python bench.py --ncol 1024 --iters 2 --steps 1 --warmup 1 --no-e2e --no-cpu"
for cfg in "$@"; do
  chunk=${cfg%%:*}; inter=${cfg##*:}
  export MALI_COL_CHUNK=$chunk MALI_DIR_INTERLEAVE=$inter
  python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('CFG $cfg fs_ms %.3f value %.4e' % (d['roofline']['mean_launch_ms'], d['value']))"
  $CMD > gpurun_out/traffic_plain.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
      --log-file gpurun_out/launches_${chunk}_${inter}.csv $CMD > gpurun_out/traffic_ncu.log 2>&1
  python tools/traffic_from_launches.py gpurun_out/launches_${chunk}_${inter}.csv gpurun_out/traffic_${chunk}_${inter}.json
done
