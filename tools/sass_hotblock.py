#!/usr/bin/env python
"""Static instruction mix of the fused kernel's hot block (no GPU needed).

    python tools/sass_hotblock.py <object-or-sass-file> [kernel-name-regex]

For every matching SASS function: the longest branch-free block (for gd_warp_kernel that
is the FAST math of one full tile = 4 rows per lane, straight-line) and its opcode
histogram.  Used to compare build variants (tools/prepare_variants.sh) before spending GPU
time: the kernel is bounded by energy per pair under the board power cap (DESIGN.md
section 4), and instructions per row is the static proxy for it."""
import collections
import re
import subprocess
import sys


def functions(path):
    if path.endswith(('.o', '.so', '.cubin')):
        txt = subprocess.run(['cuobjdump', '-sass', path], capture_output=True, text=True,
                             check=True).stdout
    else:
        txt = open(path).read()
    out = {}
    for part in re.split(r'\n\s*Function : ', txt)[1:]:
        name, body = part.split('\n', 1)
        out[name.strip()] = body
    return out


def blocks(body):
    cur, res = [], []
    for line in body.split('\n'):
        if re.match(r'\s*\.L_x_\d+:', line):
            if cur:
                res.append(cur)
                cur = []
            continue
        m = re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+(.*?);', line)
        if not m:
            continue
        ins = m.group(1)
        cur.append(ins)
        toks = ins.split()
        op = toks[1] if toks[0].startswith('@') else toks[0]
        if op.startswith(('BRA', 'EXIT', 'RET', 'CALL', 'BRX', 'JMP')) and not ins.startswith('@'):
            res.append(cur)
            cur = []
    if cur:
        res.append(cur)
    return res


def histogram(block):
    h = collections.Counter()
    for ins in block:
        toks = ins.split()
        op = toks[1] if toks[0].startswith('@') else toks[0]
        h[op.split('.')[0]] += 1
    return h


def main():
    pat = sys.argv[2] if len(sys.argv) > 2 else 'gd_warp_kernel'
    for name, body in functions(sys.argv[1]).items():
        if not re.search(pat, name):
            continue
        bs = blocks(body)
        big = max(bs, key=len)
        h = histogram(big)
        fp = sum(h[k] for k in ('FMUL', 'FFMA', 'FADD', 'FMUL2', 'FFMA2', 'FADD2'))
        print(f'{name}: {sum(len(b) for b in bs)} instructions, hot block {len(big)} '
              f'({len(big) / 4:.1f} per row), FP32 arithmetic {fp}, MUFU {h["MUFU"]}')
        print('   ' + ', '.join(f'{k} {v}' for k, v in h.most_common(14)))


if __name__ == '__main__':
    main()
