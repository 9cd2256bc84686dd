"""Build liberd_b200.so in-tree with nvcc for sm_100a (no torch, no JIT cache)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_DIR = os.path.join(HERE, 'lib')
LIB_PATH = os.environ.get('ERD_B200_LIB') or os.path.join(LIB_DIR, 'liberd_b200.so')   # (override: A/B runs of developer builds)
SOURCES = ['api.cu', 'ers.cu', 'atss.cu', 'nms.cu', 'loss.cu', 'student.cu', 'teacher.cu', 'teacher_head.cu', 'predict.cu', 'exchange.cu', 'profile.cu']


def nvcc_path() -> str:
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found')


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(os.path.dirname(HERE), 'include', 'erd_b200.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    # N ranks of a torchrun may get here at once: one builds (file lock), into a temporary file that is
    # renamed over the library atomically, so nobody ever maps a half-written .so
    import fcntl
    lock = open(LIB_PATH + '.lock', 'w')
    fcntl.flock(lock, fcntl.LOCK_EX)
    try:
        if not force and not needs_build():   # another rank built it while we waited
            return LIB_PATH
        return _build_locked(verbose)
    finally:
        fcntl.flock(lock, fcntl.LOCK_UN)
        lock.close()


def _build_locked(verbose: bool) -> str:
    tmp = LIB_PATH + f'.tmp{os.getpid()}'
    cmd = [nvcc_path(), '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
           '-shared', '-Xcompiler', '-fPIC', '--cudart', 'shared',
           '-o', tmp] + [os.path.join(CSRC, s) for s in SOURCES]
    cmd[1:1] = os.environ.get('ERD_EXTRA_NVCC', '').split()   # developer builds, e.g. -DERD_DEV_ABLATE
    if verbose:
        cmd.insert(1, '-Xptxas=-v')
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError('nvcc failed:\n' + res.stdout + res.stderr)
    os.replace(tmp, LIB_PATH)
    if verbose:
        print(res.stderr)
    return LIB_PATH


if __name__ == '__main__':
    print(build(force=True, verbose='-v' in sys.argv))
