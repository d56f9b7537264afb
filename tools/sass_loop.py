#!/usr/bin/env python3
"""Hot-path instruction estimate of a kernel's main loop from cuobjdump -sass.

    tools/sass_loop.py <object> <function-substring> [min block size to print, default 25]

Splits the kernel into basic blocks, finds the largest loop (backward branch),
and walks it taking every forward conditional branch as NOT taken unless its
target lies before the fall-through's next 'cold' marker... (heuristic): prints
the blocks of the loop in address order with size and the opcode mix, so the
steady path can be read off.  Blocks reached only through CALL.REL / slow-path
branches (the SafeMath recompute) show up as large blocks with DSETP/MUFU-heavy
mixes after a BRA.
"""
import collections
import re
import subprocess
import sys


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    min_block = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    fn = None
    ins = []
    for line in out.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            if fn and pat in fn and ins:
                break
            fn = m.group(1)
            ins = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m and fn and pat in fn:
            addr = int(m.group(1), 16)
            text = m.group(2)
            t = text.split()
            pred = t[0] if t[0].startswith("@") else ""
            op = t[1] if pred else t[0]
            tgt = None
            mt = re.search(r"0x([0-9a-f]+)", text) if op.split(".")[0] in ("BRA", "BSSY", "CALL") else None
            if mt:
                tgt = int(mt.group(1), 16)
            ins.append((addr, pred, op, tgt, text))
    if not ins:
        sys.exit("function not found")
    addrs = [a for a, *_ in ins]
    # block leaders: branch targets and instructions after branches
    leaders = {addrs[0]}
    for i, (a, pred, op, tgt, text) in enumerate(ins):
        base = op.split(".")[0]
        if base in ("BRA", "EXIT", "RET", "BSYNC", "CALL", "BREAK"):
            if i + 1 < len(ins):
                leaders.add(ins[i + 1][0])
            if tgt is not None and base == "BRA":
                leaders.add(tgt)
    # largest backward branch
    loops = [(a - tgt, tgt, a) for a, pred, op, tgt, text in ins if op.split(".")[0] == "BRA" and tgt is not None and tgt <= a]
    loops.sort(reverse=True)
    print(f"{len(ins)} instructions, loops (span bytes, head, tail):", [(s // 16, hex(h), hex(t)) for s, h, t in loops[:4]])
    if not loops:
        return
    _, head, tail = loops[0]
    blocks = []
    cur = None
    for a, pred, op, tgt, text in ins:
        if a in leaders:
            cur = [a, collections.Counter(), None]
            blocks.append(cur)
        cur[1][op.split(".")[0]] += 1
        if op.split(".")[0] in ("BRA", "EXIT", "RET", "CALL"):
            cur[2] = (pred, op, tgt)
    tot = 0
    for a, c, last in blocks:
        if a < head or a > tail:
            continue
        n = sum(c.values())
        fp64 = sum(c[k] for k in ("DFMA", "DMUL", "DADD", "DSETP"))
        mem = sum(c[k] for k in ("LDG", "STG", "LDS", "STS", "LD", "ST", "LDL", "STL"))
        lastd = f"{last[0]} {last[1]} -> {hex(last[2]) if last[2] is not None else ''}" if last else ""
        tot += n
        if n < min_block:
            continue
        print(f"  {a:#07x}: {n:5d} instr, fp64 {fp64:4d}, mem {mem:3d}, mufu {c['MUFU']:2d}  {lastd}")
    print("loop total (all blocks, incl. cold):", tot)


if __name__ == "__main__":
    main()
