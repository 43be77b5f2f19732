"""Builds ``libgdloss_b200.so`` (the C-ABI CUDA library) in-tree with nvcc.

sm_100a only (``-gencode arch=compute_100a,code=sm_100a``); nvcc cross-compiles
without a GPU.  The ``.so`` is git-ignored but travels to the GPU box with the
repo snapshot.  A content hash of the sources is stored beside the library so a
stale binary is rebuilt and a current one is not (mtimes do not survive the
snapshot copy).
"""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, 'csrc')
INCLUDE = os.path.join(os.path.dirname(PKG_DIR), 'include')
SOURCES = ['gd_loss_api.cu', 'gd_loss_inst_gwd.cu', 'gd_loss_inst_kld.cu', 'gd_loss_inst_jd.cu',
           'gd_loss_inst_symmax.cu', 'gd_loss_inst_symmin.cu', 'gd_loss_inst_bd.cu',
           'gd_loss_inst_kfiou.cu', 'gd_pairwise.cu', 'gd_host_pipeline.cu', 'gd_decoded.cu',
           'gd_pairwise_inst_gwd.cu', 'gd_pairwise_inst_kld.cu', 'gd_pairwise_inst_jd.cu',
           'gd_pairwise_inst_symmax.cu', 'gd_pairwise_inst_symmin.cu', 'gd_pairwise_inst_bd.cu',
           'gd_pairwise_inst_kfiou.cu']
HEADERS = ['gd_math.cuh', 'gd_packed.cuh', 'gd_common.cuh', 'gd_loss_kernels.cuh', 'gd_decode.cuh',
           'gd_pairwise.cuh']
LIB_NAME = 'libgdloss_b200.so'
LIB_PRECISE_NAME = 'libgdloss_b200_precise.so'   # -DGD_PRECISE_MATH=1, tests only
LIB_TUNE_NAME = 'libgdloss_b200_tune.so'         # -DGD_TUNE=1, tools/tune_sweep.py only

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo',
              '-std=c++17', '-Xcompiler', '-fPIC', '-Xcompiler', '-O3',
              '-Xcompiler', '-fvisibility=hidden', '-Xcompiler', '-fno-gnu-unique']


def _kind(precise=False, tune=False):
    return 'tune' if tune else ('precise' if precise else '')


def lib_path(precise=False, tune=False):
    name = LIB_TUNE_NAME if tune else (LIB_PRECISE_NAME if precise else LIB_NAME)
    return os.path.join(PKG_DIR, name)


def _nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found: cannot build libgdloss_b200.so')
    return exe


def _source_hash(extra=''):
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        with open(os.path.join(CSRC, name), 'rb') as f:
            h.update(name.encode())
            h.update(f.read())
    with open(os.path.join(INCLUDE, 'gd_loss_b200.h'), 'rb') as f:
        h.update(f.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    h.update(extra.encode())
    return h.hexdigest()


def is_current(precise=False, tune=False):
    lib = lib_path(precise, tune)
    stamp = lib + '.hash'
    if not (os.path.exists(lib) and os.path.exists(stamp)):
        return False
    with open(stamp) as f:
        return f.read().strip() == _source_hash(_kind(precise, tune))


def build_variant(tune_default):
    """Production-style build with the kernel knobs baked in at compile time
    (``-DGD_TUNE_DEFAULT=<bits>``): ``libgdloss_b200_v<bits>.so``, A/B measurements only."""
    lib = os.path.join(PKG_DIR, f'libgdloss_b200_v{int(tune_default)}.so')
    nvcc = _nvcc()
    build_dir = os.path.join(PKG_DIR, 'build', f'v{int(tune_default)}')
    os.makedirs(build_dir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(build_dir, src.replace('.cu', '.o'))
        res = subprocess.run([nvcc] + NVCC_FLAGS + [f'-DGD_TUNE_DEFAULT={int(tune_default)}', '-I',
                                                    INCLUDE, '-c', os.path.join(CSRC, src), '-o', obj],
                             capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src}:\n{res.stdout}\n{res.stderr}')
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    res = subprocess.run([nvcc, '-shared', '-o', lib] + objs + ['-cudart', 'static'],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f'link failed:\n{res.stdout}\n{res.stderr}')
    return lib


def build(force=False, precise=False, verbose=False, tune=False):
    """Compile the library if missing or stale; returns its path."""
    lib = lib_path(precise, tune)
    if not force and is_current(precise, tune):
        return lib
    nvcc = _nvcc()
    build_dir = os.path.join(PKG_DIR, 'build', _kind(precise, tune) or 'fast')
    os.makedirs(build_dir, exist_ok=True)
    defs = ['-DGD_TUNE=1'] if tune else (['-DGD_PRECISE_MATH=1'] if precise else [])

    def compile_one(src):
        obj = os.path.join(build_dir, src.replace('.cu', '.o'))
        cmd = [nvcc] + NVCC_FLAGS + defs + ['-I', INCLUDE, '-c',
                                            os.path.join(CSRC, src), '-o', obj]
        if verbose:
            cmd += ['-Xptxas', '-v']
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src}:\n{res.stdout}\n{res.stderr}')
        if verbose:
            sys.stderr.write(res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    res = subprocess.run([nvcc, '-shared', '-o', lib] + objs + ['-cudart', 'static'],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f'link failed:\n{res.stdout}\n{res.stderr}')
    with open(lib + '.hash', 'w') as f:
        f.write(_source_hash(_kind(precise, tune)))
    return lib


if __name__ == '__main__':
    if '--variant' in sys.argv:
        print(build_variant(int(sys.argv[sys.argv.index('--variant') + 1])))
        sys.exit(0)
    print(build(force='--force' in sys.argv, precise='--precise' in sys.argv,
                verbose='-v' in sys.argv, tune='--tune' in sys.argv))
