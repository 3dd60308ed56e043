"""Compile the C oracle (test infrastructure) into oracle/libpanoptic_oracle.so.

Recipe:  gcc -O2 -fPIC -shared -ffp-contract=off -fopenmp panoptic_oracle.c -lm
`-ffp-contract=off` is essential: the grouping distance must be evaluated as
individually rounded f32 mul/add/sub followed by ONE explicit fmaf (see the C file).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'panoptic_oracle.c')
LIB = os.path.join(HERE, 'libpanoptic_oracle.so')


def build(force: bool = False) -> str:
    if (not force and os.path.exists(LIB)
            and os.path.getmtime(LIB) >= os.path.getmtime(SRC)):
        return LIB
    cmd = ['gcc', '-O2', '-fPIC', '-shared', '-ffp-contract=off', '-fopenmp',
           '-Wall', '-Wno-unknown-pragmas', SRC, '-o', LIB, '-lm']
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv))
