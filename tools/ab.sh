#!/bin/bash
# A/B of tuning builds: tools/ab.sh <size> <tag> [<tag> ...]   ("base" = product library)
size=$1; shift
for tag in "$@"; do
  lib=""; [ "$tag" != base ] && lib="$PWD/finitevolumesolver_b200/libfvs_b200_${tag}.so"
  FVS_B200_LIB=$lib python bench.py --size $size --steps 5 --no-cpu-baseline --e2e-steps 1 --no-secondary --no-strong --no-parity-check 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); o=d['other_arithmetic']; print('$tag', 'G/s %.3f'%(d['value']/1e9), 'step ms %.2f'%d['ms_per_step'], 'sweep ms %.3f'%d['roofline']['avg_launch_ms'], 'frac %.4f'%d['roofline']['frac'], '| other:', 'G/s %.3f'%(o['value']/1e9), 'sweep ms %.3f'%o['avg_launch_ms'], 'frac %.4f'%o['roofline_frac'])"
done
