#!/bin/bash
# first GPU call of round 2: full -m gpu suite, compute-sanitizer summaries, baseline bench of the round-1 kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2a_pytest.log
for tool in memcheck racecheck; do
  for fx in c1_falc_ca c2_falc_cah; do
    timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py $fx > gpurun_out/r2a_san_${tool}_${fx}.txt 2>&1
  done
  MALI_NO_SPEC=1 timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py c1_falc_ca > gpurun_out/r2a_san_${tool}_c1_nospec.txt 2>&1
done
python bench.py --steps 3 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2a_bench_ref.json 2>> gpurun_out/r2a_bench.err
tail -3 gpurun_out/r2a_pytest.log
for f in gpurun_out/r2a_san_*.txt; do echo $f; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_run" $f | tail -3; done
