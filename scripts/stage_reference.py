#!/usr/bin/env python
"""Install the UNMODIFIED reference into the git-ignored `baseline/_ref/` (authoring container
only: `/root/reference` does not exist on the GPU box, `baseline/_ref/` travels there with the
gpurun snapshot).

    python scripts/stage_reference.py [--force]

`pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of
/root/reference>`: the source tree is read-only, so it is built from a copy under /tmp; --no-deps
because `torchmetrics` and `nicr_scene_analysis_datasets` are in neither the image nor the
wheelhouse (stand-ins: oracle/ref_stubs, used for state bookkeeping / import hooks only).
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOURCE = '/root/reference'
TARGET = os.path.join(ROOT, 'baseline', '_ref')


def installed() -> bool:
    return os.path.isdir(os.path.join(TARGET, 'nicr_mt_scene_analysis'))


def stage(force: bool = False) -> str:
    """-> 'installed' | 'present' | 'no source'"""
    if installed() and not force:
        return 'present'
    if not os.path.isdir(os.path.join(SOURCE, 'src', 'nicr_mt_scene_analysis')):
        return 'no source'
    tmp = tempfile.mkdtemp(prefix='npb_ref_')
    try:
        copy = os.path.join(tmp, 'reference')
        shutil.copytree(SOURCE, copy)
        shutil.rmtree(TARGET, ignore_errors=True)
        subprocess.run([sys.executable, '-m', 'pip', 'install', '--no-index',
                        '--no-build-isolation', '--no-deps', '--find-links', '/opt/wheelhouse',
                        '--target', TARGET, copy], check=True, capture_output=True)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return 'installed' if installed() else 'no source'


if __name__ == '__main__':
    print('baseline/_ref:', stage(force='--force' in sys.argv))
