"""Developer helper: instruction mix of the DFMA-carrying loops of one kernel (cuobjdump -sass text on stdin)."""
import re
import sys
from collections import Counter

ops = []
addr = {}
for l in sys.stdin:
    m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*);', l)
    if not m:
        continue
    a = int(m.group(1), 16)
    addr[a] = len(ops)
    ops.append((a, m.group(2)))
thr = int(sys.argv[1]) if len(sys.argv) > 1 else 50
for i, (a, ins) in enumerate(ops):
    m = re.search(r'BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)', ins)
    if m:
        t = int(m.group(1), 16)
        if t < a and t in addr:
            body = [o for _, o in ops[addr[t]:i + 1]]
            nd = sum('DFMA' in o for o in body)
            if nd > thr:
                c = Counter(re.sub(r'^@!?U?P\d+\s+', '', o).split()[0].split('.')[0] for o in body)
                print(f'loop {t:x}-{a:x}: {len(body)} instrs, {nd} DFMA ({100*nd/len(body):.0f}%)', dict(c.most_common(12)))
