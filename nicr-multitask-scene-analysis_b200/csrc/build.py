"""Build libnicr_panoptic_b200.so (hand-written sm_100a CUDA behind a C ABI) in-tree.

    nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -shared
         -Xcompiler -fPIC -I include  csrc/*.cu  -o csrc/libnicr_panoptic_b200.so

No fast-math: the grouping distance relies on IEEE sqrtf / explicit rounding intrinsics.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, 'libnicr_panoptic_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')


def sources():
    return sorted(glob.glob(os.path.join(HERE, '*.cu')))


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    deps = sources() + glob.glob(os.path.join(HERE, '*.cuh')) + \
        glob.glob(os.path.join(ROOT, 'include', '*.h'))
    return any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    cmd = [NVCC, '-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
           '-shared', '-Xcompiler', '-fPIC', '-I', os.path.join(ROOT, 'include'), '-I', HERE]
    if verbose:
        cmd += ['-Xptxas', '-v']
    cmd += sources() + ['-o', LIB]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
