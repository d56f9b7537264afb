#!/bin/bash
# counts the instructions of the longest straight-line block (the steady-state
# body) of the 3-D MUSCL-Hancock<HLLEM-L, VanLeer> z-sweep kernel for a given
# set of -D flags:  tools/count_steady.sh [-DNAME=VALUE ...]
cd "$(dirname "$0")/../finitevolumesolver_b200/csrc"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=${FMAD:-false} -ccbin /usr/bin/g++ -I ../../include -DFVS_INST_D=3 -DFVS_INST_BASE=0 "$@" -c sweep_inst.cu -o /tmp/cs.o 2>/dev/null || exit 1
for fn in _ZN3fvs26fused_sweep_yz_smem_kernelILi3ELi2ELi0ELi1ELi2EEEvNS_9SweepArgsE _ZN3fvs25fused_sweep_x_smem_kernelILi3ELi0ELi1ELi2EEEvNS_9SweepArgsE; do
cuobjdump -sass -fun $fn /tmp/cs.o | grep -E '^\s+/\*[0-9a-f]{4,5}\*/' > /tmp/cs.sass
python3 - <<'PY'
import re,collections
ops=[]
for l in open('/tmp/cs.sass'):
    m=re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);',l)
    if not m: continue
    t=m.group(2).split()
    ops.append(t[1] if t[0].startswith('@') else t[0])
runs=[];start=0
for i,op in enumerate(ops):
    if op.split('.')[0] in ('BRA','CALL','BSSY','BSYNC','EXIT','RET','BAR'):
        runs.append((i-start,start,i)); start=i+1
runs.sort(reverse=True)
n,s,e=runs[0]
c=collections.Counter(o.split('.')[0] for o in ops[s:e])
fp64=sum(c[k] for k in ('DFMA','DMUL','DADD','DSETP'))
print(f"total {len(ops)} | longest block {n}: fp64 {fp64}, other {n-fp64} |", dict(c.most_common(12)))
PY
done
