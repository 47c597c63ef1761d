"""Development aid: print the non-core instructions of the main recurrence loop of /tmp/k<K>.sass."""
import re, sys
K = sys.argv[1]
show_all = len(sys.argv) > 2
ins = []
for l in open(f'/tmp/k{K}.sass'):
    m = re.match(r'\s*/\*([0-9a-f]{4,5})\*/\s*(.*?);', l)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
addr = {a: i for i, (a, _) in enumerate(ins)}
for i, (a, t) in enumerate(ins):
    m = re.search(r'BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)', t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < a and tgt in addr:
            body = ins[addr[tgt]:i + 1]
            n = sum(1 for _, x in body if 'SHFL.UP' in x)
            if n in (4, 8):
                for a2, x in body:
                    if show_all or not re.match(r'(FADD|FADD2|FMNMX|SHF\.L|LDS R\d+, \[R\d+\+UR)', x): print(hex(a2), x)
                break
