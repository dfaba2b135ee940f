"""TEST INFRASTRUCTURE ONLY -- compile the C restatements under oracle/ with gcc (no GPU involved).

    python -m oracle.build          # -> oracle/_build/libphysics_port.so

Called by __graft_entry__.build(); the .so is git-ignored but travels to the GPU box with the snapshot."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, '_build')
PHYSICS_LIB = os.path.join(OUT_DIR, 'libphysics_port.so')


def build(force: bool = False) -> str:
    src = os.path.join(HERE, 'physics_port.c')
    if not force and os.path.exists(PHYSICS_LIB) and os.path.getmtime(PHYSICS_LIB) >= os.path.getmtime(src):
        return PHYSICS_LIB
    gcc = shutil.which('gcc')
    if gcc is None:
        raise RuntimeError('gcc not found: cannot build the C oracle')
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = [gcc, '-O3', '-fopenmp', '-shared', '-fPIC', '-o', PHYSICS_LIB, src, '-lm']
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('gcc failed:\n' + res.stdout + res.stderr)
    return PHYSICS_LIB


if __name__ == '__main__':
    print(build(force=True))
