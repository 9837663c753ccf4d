#!/bin/bash
# ncu --set full of the kernels around the formal solution (finish kernels, statistical equilibrium) and of the upload
# path (re-layout, line profiles) on a 512-column chunk, after the same command ran clean without ncu
mkdir -p gpurun_out
tag=${1:-v23}
CMD="python bench.py --ncol 512 --iters 2 --steps 1 --warmup 1 --no-cpu --e2e-steps 1"
$CMD > gpurun_out/ncu_small_${tag}_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:"compute_phi_kernel|pack_tiles_kernel|j_finish_kernel|gamma_finish_kernel|stat_equil_kernel|cont_group_kernel" \
    -s 18 -c 8 -o /tmp/prof_small_${tag} -f $CMD > gpurun_out/ncu_small_${tag}.log 2>&1
tail -2 gpurun_out/ncu_small_${tag}.log
python tools/ncu_summary.py /tmp/prof_small_${tag}.ncu-rep > gpurun_out/ncu_small_${tag}_summary.txt 2>&1
grep -E "Kernel Name|time_duration|dram_throughput|l1tex__data_pipe_lsu_wavefronts.avg|fp64_cycles|issue_active" gpurun_out/ncu_small_${tag}_summary.txt | head -60
