#!/bin/bash
# one ncu --set full capture of the three formal-solution kernels (after the same command ran clean without ncu); the
# report of these very large kernels exceeds what gpurun copies back, so it is summarised on the box
# (tools/ncu_summary.py: key metrics, stall reasons, opcode mix) and only the summaries travel
mkdir -p gpurun_out
tag=${1:-fs}
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --ncol 1024 --iters 1"
MALI_SERIAL_CLASSES=1 $CMD > gpurun_out/ncu_${tag}_plain.log 2>&1 && \
MALI_SERIAL_CLASSES=1 ncu --set full --clock-control none -k regex:fs_gamma_kernel_m -s 3 -c 3 \
    -o /tmp/prof_${tag} -f $CMD > gpurun_out/ncu_${tag}.log 2>&1
tail -2 gpurun_out/ncu_${tag}.log
python tools/ncu_summary.py /tmp/prof_${tag}.ncu-rep > gpurun_out/ncu_${tag}_summary.txt 2>&1
ncu -i /tmp/prof_${tag}.ncu-rep --page raw --csv > gpurun_out/ncu_${tag}_raw.csv 2>/dev/null
ls -la /tmp/prof_${tag}.ncu-rep gpurun_out/
