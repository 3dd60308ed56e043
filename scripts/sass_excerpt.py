#!/usr/bin/env python
"""profiles/<round>_sass_excerpt.txt: `cuobjdump -sass` of the hot kernels of the in-tree library --
instruction count, mnemonic histogram and the instructions that characterise each kernel
(128-bit streaming loads / stores, packed FP32, MATCH / REDUX, shared and global atomics).

    python scripts/sass_excerpt.py [r02]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'nicr-multitask-scene-analysis_b200', 'csrc', 'libnicr_panoptic_b200.so')
KERNELS = [
    '_ZN3npb19group_pixels_kernelILi4ELi0ELb0ELi64EEEvNS_11GroupParamsE',
    '_ZN3npb19group_pixels_kernelILi4ELi0ELb1ELi256EEEvNS_11GroupParamsE',
    '_ZN3npb17pair_count_kernelILi4ELb1ELb1ELb1EEEvNS_10PairParamsE',
    '_ZN3npb17pair_count_kernelILi4ELb1ELb1ELb0EEEvNS_10PairParamsE',
    '_ZN3npb21confmat_stream_kernelIxhLb0EEEvPKT_PKT0_xiPyPi',
    '_ZN3npb19match_frames_kernelENS_11MatchParamsE',
    '_ZN3npb28nms_candidates_direct_kernelILi4EEEvPKfiifiP5uint2iPiNS_12SelectParamsE',
]
SELECT = re.compile(r'LDG\.E\.EF\.128|STG\.E\.EF\.128|F(ADD|MUL|FMA)2|FMNMX\.NAN|MATCH|REDUX|ATOMS|'
                    r'REDG|ATOMG|UTMA|UTC|LDTM')


def main():
    rnd = sys.argv[1] if len(sys.argv) > 1 else 'r02'
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    blocks = {}
    name = None
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            name = m.group(1)
            blocks[name] = []
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(.*?);', line)
        if m and name:
            blocks[name].append(m.group(1).strip())
    regs = subprocess.run(['cuobjdump', '-res-usage', LIB], capture_output=True, text=True).stdout
    usage = dict(re.findall(r'Function (\S+):\n\s+(REG:\d+ STACK:\d+ SHARED:\d+)', regs))
    out = ['# cuobjdump -sass excerpts of libnicr_panoptic_b200.so (sm_100a), round ' + rnd[1:].lstrip('0'),
           '']
    for k in KERNELS:
        ins = blocks.get(k)
        if not ins:
            continue
        out.append('## ' + k)
        out.append(f'instructions: {len(ins)}   {usage.get(k, "")}')
        hist = collections.Counter((i.split()[1] if i.startswith('@') else i.split()[0]) for i in ins)
        out.append('mnemonic histogram (top 24):')
        out += [f'{n:7d} {m}' for m, n in hist.most_common(24)]
        out.append('selected instructions:')
        sel = collections.Counter(re.sub(r'R\d+|UR\d+|P\d', '_', i) for i in ins if SELECT.search(i))
        out += [f'{n:7d}    {i} ;' for i, n in sel.most_common(16)]
        out.append('')
    path = os.path.join(ROOT, 'profiles', f'{rnd}_sass_excerpt.txt')
    open(path, 'w').write('\n'.join(out))
    print(path, len(out), 'lines')


if __name__ == '__main__':
    main()
