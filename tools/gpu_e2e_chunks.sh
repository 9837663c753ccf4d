#!/bin/bash
# e2e pipeline experiment: chunking of the host->device upload
for cfg in "512:0" "512:128" "448:128" "256:0" "256:64" "384:128"; do
  chunk=${cfg%%:*}; first=${cfg##*:}
  python bench.py --steps 3 --warmup 2 --no-cpu --chunk $chunk --first-chunk $first 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('CHUNK $cfg value %.4e e2e %.4e (%.1f ms/step) host_phi %.4e' % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e_host_phi']['value']))"
done
