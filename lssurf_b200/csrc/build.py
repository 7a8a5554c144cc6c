"""Build liblssurf_b200.so in-tree with nvcc for sm_100a (B200).

    python -m lssurf_b200.csrc.build [--force]

nvcc cross-compiles without a GPU.  The shared library lands next to the package
(`lssurf_b200/liblssurf_b200.so`), is git-ignored, and travels to the GPU box with the snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
SOURCES = ['assembly.cu', 'sort.cu', 'spmv.cu', 'comm.cu', 'normal.cu', 'precond.cu', 'coarse.cu', 'structured.cu', 'solver.cu', 'robust.cu']
LIB = os.path.join(PKG, 'liblssurf_b200.so')
NVCC_FLAGS = ['-O3', '-std=c++17', '-lineinfo', '-gencode', 'arch=compute_100a,code=sm_100a',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=default']


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(HERE, f) for f in SOURCES + ['common.cuh', 'precond.cuh', 'coarse.cuh', 'structured.cuh']] + \
           [os.path.join(os.path.dirname(PKG), 'include', 'lssurf_b200.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=True):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(HERE, src.replace('.cu', '.o'))
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + ['-c', os.path.join(HERE, src), '-o', obj]
        if verbose:
            print(' '.join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if out.strip() and verbose:
            print(out)
        if p.returncode != 0:
            failed = True
            print(f'nvcc failed on {src}:\n{out}', file=sys.stderr)
    if failed:
        raise RuntimeError('nvcc compilation failed')
    cmd = [nvcc, '-shared', '-o', LIB] + objs + ['-ldl']
    if verbose:
        print(' '.join(cmd), flush=True)
    subprocess.check_call(cmd)
    return LIB


if __name__ == '__main__':
    build(force='--force' in sys.argv)
    print(LIB)
