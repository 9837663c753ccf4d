#!/bin/bash
# quick GPU check used during kernel tuning: parity tests (fast subset) + resident bench
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "single_formal or c1_free or determin" 2>&1 | tail -3
timeout 300 python bench.py --no-e2e --no-cpu "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.3e  ms/step %.1f  fs_ms %.2f  frac %.4f'%(d['value'], d['ms_per_step'], d['roofline']['mean_launch_ms'], d['roofline']['frac']))"
