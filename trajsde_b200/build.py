"""Build the in-tree CUDA library: nvcc -> trajsde_b200/lib/libtrajsde_b200.so (sm_100a only, cross-compiles without a GPU).

    python -m trajsde_b200.build [--force] [--verbose]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_DIR = os.path.join(HERE, 'lib')
LIB_PATH = os.path.join(LIB_DIR, 'libtrajsde_b200.so')
SOURCES = ['abi.cu', 'euler_exact.cu', 'euler_bwd_exact.cu', 'euler_bwd_tc.cu', 'gru_bwd.cu', 'gru_bwd_tc.cu', 'enc_bwd.cu', 'enc_bwd_sweep.cu', 'euler_tc.cu', 'enc_tc.cu', 'heads.cu', 'heads_bwd.cu', 'stage_ops.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '--use_fast_math=false',
              '-Xcompiler', '-fPIC', '-Xptxas', '-v']


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isfile(cand) or cand == 'nvcc'):
            return cand
    raise RuntimeError('nvcc not found')


def _stale():
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, '..', 'include', 'trajsde_b200.h'), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIB_DIR, src.replace('.cu', '.o'))
        cmd = [_nvcc()] + [f for f in NVCC_FLAGS if f != '--use_fast_math=false'] + ['-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((src, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    for src, cmd, p in procs:
        out, _ = p.communicate()
        log.append(f'== {src}\n{out}')
        if p.returncode != 0:
            raise RuntimeError(f'nvcc failed for {src}:\n{" ".join(cmd)}\n{out}')
    cmd = [_nvcc(), '-shared', '-o', LIB_PATH] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a']
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}')
    with open(os.path.join(LIB_DIR, 'build.log'), 'w') as f:
        f.write('\n'.join(log))
    if verbose:
        print('\n'.join(log))
    return LIB_PATH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
