"""Build libnicr_panoptic_b200.so (hand-written sm_100a CUDA behind a C ABI) in-tree.

    nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -shared -cudart shared
         -Xcompiler -fPIC -I include  csrc/*.cu  -o csrc/libnicr_panoptic_b200.so

(the CUDA runtime is linked dynamically: the process already holds torch's libcudart, and the
library stays free of the runtime's own entry points)

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


def build(force: bool = False, verbose: bool = False, defines=(), out: str = LIB) -> str:
    """`defines` / `out`: instrumented variants of the same ABI for measurements and debugging
    (loaded through NPB_LIB_PATH), e.g. build(force=True, defines=('NPB_TIMELINE',),
    out='build/timeline/libnicr_panoptic_b200.so')."""
    if not force and out == LIB and not is_stale():
        return LIB
    os.makedirs(os.path.dirname(os.path.abspath(out)), exist_ok=True)
    cmd = [NVCC, '-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
           '-shared', '-cudart', 'shared', '-Xcompiler', '-fPIC', '-I', os.path.join(ROOT, 'include'),
           '-I', HERE]
    cmd += ['-D' + d for d in defines]
    if verbose:
        cmd += ['-Xptxas', '-v']
    cmd += sources() + ['-o', out]
    subprocess.run(cmd, check=True)
    return out


if __name__ == '__main__':
    defs = tuple(a[2:] for a in sys.argv[1:] if a.startswith('-D'))
    outs = [a[6:] for a in sys.argv[1:] if a.startswith('--out=')]
    print(build(force='--force' in sys.argv or bool(defs), verbose='-v' in sys.argv, defines=defs,
                out=outs[0] if outs else LIB))
