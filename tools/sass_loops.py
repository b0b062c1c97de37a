#!/usr/bin/env python
"""Static view of the hot loops of a kernel: every backward branch of its SASS with the instruction mix of the body.

    python tools/sass_loops.py <object or .so> <substring of the kernel name> [--dump]

Reads `cuobjdump -sass`; a loop is [target of a backward BRA, the BRA].  Nested loops are reported separately (the
inner body is part of the outer one's count).  Used to budget instructions per segment-step before spending GPU time
(DESIGN.md 3.3); --dump prints the loop bodies (the excerpts committed under profiles/)."""
import re
import subprocess
import sys
from collections import Counter

CLASSES = [('fp64', r'^(DFMA|DMUL|DADD|DSETP|DMNMX)'), ('mufu', r'^MUFU'), ('lds', r'^LDS'), ('sts', r'^STS'),
           ('ldg', r'^(LDG|LD\.)'), ('stg', r'^(STG|ST\.)'), ('local', r'^(LDL|STL)'), ('ldgsts', r'^(LDGSTS|LDGDEPBAR|DEPBAR)'),
           ('bar', r'^(BAR|WARPSYNC|BSSY|BSYNC|VOTE)'), ('branch', r'^(BRA|BRX|EXIT|RET|CALL)'),
           ('int', r'^(IMAD|IADD|IADD3|LOP3|LEA|SHF|VIADD|VIMNMX|ISETP|SEL|PRMT|MOV|IABS|SGXT|BMSK|PLOP3|UMOV|UIADD3|ULOP3|UIMAD|ULEA|USHF|S2R|S2UR|R2UR|CS2R|UISETP|USEL|UPLOP3|FSEL|POPC|FLO|I2F|F2I|F2F|R2P|P2R|LDC|ULDC|LDCU|NOP|HFMA2|FFMA|FMUL|FADD|FSETP|FMNMX)')]


def sass(obj, name):
    out = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True, check=True).stdout
    funcs, cur, key = {}, None, None
    for line in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            key = m.group(1)
            cur = funcs.setdefault(key, [])
            continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
        if m and cur is not None:
            cur.append((int(m.group(1), 16), m.group(2).strip()))
    hits = [k for k in funcs if name in k]
    if len(hits) != 1:
        raise SystemExit('kernel name matches {}: {}'.format(len(hits), hits[:8]))
    return hits[0], funcs[hits[0]]


def opcode(text):
    t = re.sub(r'^@!?U?P\d+\s+', '', text)
    return t.split()[0]


def classify(op):
    base = op.split('.')[0]
    for cls, pat in CLASSES:
        if re.match(pat, base) or re.match(pat, op):
            return cls
    return 'other:' + base


def main():
    obj, name = sys.argv[1], sys.argv[2]
    dump = '--dump' in sys.argv
    fn, ins = sass(obj, name)
    addr_index = {a: i for i, (a, _) in enumerate(ins)}
    print('{}: {} instructions'.format(fn, len(ins)))
    loops = []
    for i, (a, t) in enumerate(ins):
        op = opcode(t)
        if op.startswith('BRA'):
            m = re.search(r'`\(\.L_x_\d+\)|0x([0-9a-f]+)', t)
            tgt = re.findall(r'0x([0-9a-f]+)', t)
            if tgt:
                ta = int(tgt[-1], 16)
                if ta <= a and ta in addr_index:
                    loops.append((addr_index[ta], i))
    for (s, e) in sorted(loops):
        body = ins[s:e + 1]
        ops = [opcode(t) for _, t in body]
        cls = Counter(classify(o) for o in ops)
        print('loop 0x{:04x}..0x{:04x}: {:4d} instr  '.format(ins[s][0], ins[e][0], len(body)) +
              ' '.join('{}={}'.format(k, v) for k, v in sorted(cls.items(), key=lambda kv: -kv[1])))
        if dump:
            for a, t in body:
                print('    /*{:04x}*/ {}'.format(a, t))


if __name__ == '__main__':
    main()
