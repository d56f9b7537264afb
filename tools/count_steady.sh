#!/bin/bash
# Instruction statistics of the bench kernels (3-D MUSCL-Hancock<HLLEM-L, VanLeer>
# y/z and x sweeps) for a given set of -D flags:  tools/count_steady.sh [-DNAME=VALUE ...]
# Prints, per kernel: registers / spills, total instructions, and the opcode mix of
# the longest straight-line block (the steady-state body).
cd "$(dirname "$0")/../finitevolumesolver_b200/csrc"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=${FMAD:-false} -ccbin /usr/bin/g++ -I ../../include -DFVS_INST_D=3 -DFVS_INST_BASE=0 -DFVS_INST_MINIMAL -Xptxas -v "$@" -c sweep_inst.cu -o /tmp/cs.o 2>/tmp/cs.log || { grep -i error /tmp/cs.log; exit 1; }
grep -A2 "Function properties" /tmp/cs.log | grep -v "^--" | paste - - - | grep fused | sed 's/.*fused_sweep_\([a-z_]*\)_kernelILi3E\(Li[0-9]E\).*stack frame, \([0-9]* bytes spill stores\).*Used \([0-9]* registers\).*, \([0-9]* bytes smem\)/\1 \2: \4, \3, \5/'
for fn in $(cuobjdump -sass /tmp/cs.o | grep -o "Function : .*fused_sweep[^ ]*" | sed 's/Function : //'); do
cuobjdump -sass -fun $fn /tmp/cs.o | grep -E '^\s+/\*[0-9a-f]{4,5}\*/' > /tmp/cs.sass
python3 - "$fn" <<'PY'
import re,collections,sys
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
name=re.sub(r'.*fused_sweep_(\w+?)_kernelILi3E(Li\dE).*',r'\1 \2',sys.argv[1])
print(f"{name}: total {len(ops)} | longest block {n}: fp64 {fp64}, other {n-fp64} |", dict(c.most_common(14)))
PY
done
