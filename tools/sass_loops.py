#!/usr/bin/env python3
"""Summarise SASS per kernel: instruction histogram of every backward-branch loop body.
usage: sass_loops.py <binary|.so|.o> [kernel-name-substring]"""
import re, subprocess, sys, collections
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
sel = sys.argv[2] if len(sys.argv) > 2 else ""
fn = None; ins = {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m: fn = m.group(1); ins[fn] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m and fn: ins[fn].append((int(m.group(1), 16), m.group(2).strip()))
for fn, lst in ins.items():
    if sel not in fn: continue
    print("==", fn, len(lst), "instructions")
    addr2i = {a: i for i, (a, _) in enumerate(lst)}
    for i, (a, t) in enumerate(lst):
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr2i:
                body = lst[addr2i[tgt]:i + 1]
                h = collections.Counter()
                for _, x in body:
                    x = re.sub(r"^@!?U?P\d\s+", "", x)
                    h[x.split()[0].split(".")[0]] += 1
                print(f"  loop {tgt:#x}..{a:#x}: {len(body)} instrs:", dict(h.most_common()))
