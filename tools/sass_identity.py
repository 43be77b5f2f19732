#!/usr/bin/env python
"""Compare two `cuobjdump -sass` dumps function by function (whitespace, encodings and the
trailing PACK template flag of gd_warp_kernel ignored; functions that only differ in
predicate-register numbering are reported separately).  Used to prove that a refactor of the
shared math headers leaves the already validated kernels untouched:

    python tools/sass_identity.py before.sass after.sass
"""
import re
import sys


def funcs(fn):
    out, cur = {}, None
    for line in open(fn):
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = m.group(1)
            out[cur] = []
            continue
        if cur is not None:
            s = re.sub(r'/\* 0x[0-9a-f]+ \*/', '', line).strip()
            if s:
                out[cur].append(re.sub(r'\s+', ' ', s))
    return out


def norm(name):
    if 'gd_warp_kernel' in name:
        return re.sub(r'(ELin?\d+)ELb0EEEvNS_8LossArgsE$', r'\1EEEvNS_8LossArgsE', name)
    return name


def main():
    b, a = funcs(sys.argv[1]), funcs(sys.argv[2])
    amap = {norm(k): v for k, v in a.items()}
    same = pred = diff = missing = 0
    for k, v in b.items():
        v2 = amap.get(norm(k))
        if v2 is None:
            missing += 1
            print('missing in after:', k)
        elif v == v2:
            same += 1
        elif [re.sub(r'U?P\d', 'P#', x) for x in v] == [re.sub(r'U?P\d', 'P#', x) for x in v2]:
            pred += 1
        else:
            diff += 1
            print('DIFFERENT:', k, len(v), len(v2))
    new = len(a) - (same + pred + diff)
    print(f'{sys.argv[1]}: identical {same}, predicate-renamed {pred}, different {diff}, '
          f'missing {missing}, new functions {new}')
    return 1 if (diff or missing) else 0


if __name__ == '__main__':
    sys.exit(main())
