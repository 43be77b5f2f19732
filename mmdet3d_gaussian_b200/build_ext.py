"""Builds the native code in-tree: ``libgdloss_b200.so`` (the C-ABI CUDA library, nvcc) and
``_C.so`` (the torch C++ extension over that ABI, g++).

sm_100a only (``-gencode arch=compute_100a,code=sm_100a``); nvcc cross-compiles
without a GPU.  The binaries are git-ignored but travel to the GPU box with the
repo snapshot.  A content hash of the sources is stored beside each binary so a
stale one is rebuilt and a current one is not (mtimes do not survive the
snapshot copy).

Concurrency: every build takes an inter-process lock (``fcntl.flock`` on
``.build.lock``), compiles into a per-process temporary directory and installs the
result with an atomic ``os.replace`` -- N ranks started by torchrun on a fresh
checkout build once and never load a half-written file.
"""
import contextlib
import fcntl
import hashlib
import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, 'csrc')
INCLUDE = os.path.join(os.path.dirname(PKG_DIR), 'include')
SOURCES = ['gd_loss_api.cu', 'gd_loss_inst_gwd.cu', 'gd_loss_inst_kld.cu', 'gd_loss_inst_jd.cu',
           'gd_loss_inst_symmax.cu', 'gd_loss_inst_symmin.cu', 'gd_loss_inst_bd.cu',
           'gd_loss_inst_kfiou.cu', 'gd_pairwise.cu', 'gd_host_pipeline.cu', 'gd_decoded.cu',
           'gd_pairwise_inst_gwd.cu', 'gd_pairwise_inst_kld.cu', 'gd_pairwise_inst_jd.cu',
           'gd_pairwise_inst_symmax.cu', 'gd_pairwise_inst_symmin.cu', 'gd_pairwise_inst_bd.cu',
           'gd_pairwise_inst_kfiou.cu', 'gd_simota.cu', 'gd_symm.cu']
HEADERS = ['gd_math.cuh', 'gd_packed.cuh', 'gd_common.cuh', 'gd_loss_kernels.cuh', 'gd_decode.cuh',
           'gd_pairwise.cuh']
SOURCES = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
LIB_NAME = 'libgdloss_b200.so'
LIB_PRECISE_NAME = 'libgdloss_b200_precise.so'   # -DGD_PRECISE_MATH=1, tests only
LIB_TUNE_NAME = 'libgdloss_b200_tune.so'         # -DGD_TUNE=1, tools/tune_sweep.py only
SHIM_NAME = '_C.so'                              # torch C++ extension (csrc/torch_shim.cpp)
SHIM_SOURCE = 'torch_shim.cpp'

# Compile-time default of the fused kernel's knobs (csrc/gd_loss_kernels.cuh, kTune*):
# 256 min/max row screen + 512 default alpha / center_offset folded + 1024 packed-FP32 math.
# Chosen from the interleaved A/B of round 2 (profiles/r02a_ab_variants.md): parity-green
# and +2.2 % over the round-1 kernel on the four bench configurations.
TUNE_DEFAULT = 1792

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo',
              '-std=c++17', '-Xcompiler', '-fPIC', '-Xcompiler', '-O3',
              '-Xcompiler', '-fvisibility=hidden', '-Xcompiler', '-fno-gnu-unique']


def _kind(precise=False, tune=False):
    return 'tune' if tune else ('precise' if precise else '')


def lib_path(precise=False, tune=False):
    name = LIB_TUNE_NAME if tune else (LIB_PRECISE_NAME if precise else LIB_NAME)
    return os.path.join(PKG_DIR, name)


def shim_path():
    return os.path.join(PKG_DIR, SHIM_NAME)


def _nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found: cannot build libgdloss_b200.so')
    return exe


@contextlib.contextmanager
def _build_lock():
    fd = os.open(os.path.join(PKG_DIR, '.build.lock'), os.O_CREAT | os.O_RDWR, 0o644)
    try:
        fcntl.flock(fd, fcntl.LOCK_EX)
        yield
    finally:
        fcntl.flock(fd, fcntl.LOCK_UN)
        os.close(fd)


def _defs(precise, tune, tune_default=None):
    if tune:
        return ['-DGD_TUNE=1']
    defs = [f'-DGD_TUNE_DEFAULT={TUNE_DEFAULT if tune_default is None else int(tune_default)}']
    if precise:
        defs.append('-DGD_PRECISE_MATH=1')
    return defs


def _source_hash(precise=False, tune=False):
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        with open(os.path.join(CSRC, name), 'rb') as f:
            h.update(name.encode())
            h.update(f.read())
    with open(os.path.join(INCLUDE, 'gd_loss_b200.h'), 'rb') as f:
        h.update(f.read())
    h.update(' '.join(NVCC_FLAGS + _defs(precise, tune)).encode())
    return h.hexdigest()


def _is_current(path, digest):
    stamp = path + '.hash'
    if not (os.path.exists(path) and os.path.exists(stamp)):
        return False
    with open(stamp) as f:
        return f.read().strip() == digest


def is_current(precise=False, tune=False):
    return _is_current(lib_path(precise, tune), _source_hash(precise, tune))


def _install(tmp_file, final_path, digest=None):
    """Atomic: readers see either the old complete file or the new complete file."""
    os.replace(tmp_file, final_path)
    if digest is not None:
        fd, tmp = tempfile.mkstemp(dir=PKG_DIR, suffix='.hash.tmp')
        with os.fdopen(fd, 'w') as f:
            f.write(digest)
        os.replace(tmp, final_path + '.hash')


def _compile_library(lib, defs, verbose=False):
    nvcc = _nvcc()
    build_dir = tempfile.mkdtemp(prefix=f'pid{os.getpid()}_', dir=_build_root())
    try:
        def compile_one(src):
            obj = os.path.join(build_dir, src.replace('.cu', '.o'))
            cmd = [nvcc] + NVCC_FLAGS + defs + ['-I', INCLUDE, '-c', os.path.join(CSRC, src),
                                                '-o', obj]
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
        out = os.path.join(build_dir, os.path.basename(lib))
        res = subprocess.run([nvcc, '-shared', '-o', out] + objs + ['-cudart', 'static'],
                             capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f'link failed:\n{res.stdout}\n{res.stderr}')
        return out, build_dir
    except Exception:
        shutil.rmtree(build_dir, ignore_errors=True)
        raise


def _build_root():
    root = os.path.join(PKG_DIR, 'build')
    os.makedirs(root, exist_ok=True)
    return root


def build_variant(tune_default):
    """Production-style build with other kernel knobs baked in at compile time
    (``-DGD_TUNE_DEFAULT=<bits>``): ``libgdloss_b200_v<bits>.so``, A/B measurements only."""
    lib = os.path.join(PKG_DIR, f'libgdloss_b200_v{int(tune_default)}.so')
    with _build_lock():
        out, build_dir = _compile_library(lib, _defs(False, False, tune_default))
        _install(out, lib)
        shutil.rmtree(build_dir, ignore_errors=True)
    return lib


def build(force=False, precise=False, verbose=False, tune=False):
    """Compile the library if missing or stale; returns its path."""
    lib = lib_path(precise, tune)
    digest = _source_hash(precise, tune)
    if not force and _is_current(lib, digest):
        return lib
    with _build_lock():
        if not force and _is_current(lib, digest):     # another process built it meanwhile
            return lib
        out, build_dir = _compile_library(lib, _defs(precise, tune), verbose)
        _install(out, lib, digest)
        shutil.rmtree(build_dir, ignore_errors=True)
    return lib


# ---------------------------------------------------------------------------
# torch C++ extension (the shim): g++ against the torch headers of THIS interpreter
# ---------------------------------------------------------------------------
def _shim_hash():
    import torch
    h = hashlib.sha256()
    with open(os.path.join(CSRC, SHIM_SOURCE), 'rb') as f:
        h.update(f.read())
    with open(os.path.join(INCLUDE, 'gd_loss_b200.h'), 'rb') as f:
        h.update(f.read())
    h.update(torch.__version__.encode())
    h.update(sys.version.encode())
    return h.hexdigest()


def shim_is_current():
    return _is_current(shim_path(), _shim_hash())


def build_shim(force=False):
    """Compile ``csrc/torch_shim.cpp`` into ``_C.so`` if missing or stale; returns its path."""
    out_path = shim_path()
    digest = _shim_hash()
    if not force and _is_current(out_path, digest):
        return out_path
    import torch
    from torch.utils import cpp_extension as ce
    gxx = shutil.which('g++')
    if gxx is None:
        raise RuntimeError('g++ not found: cannot build the torch shim (_C.so)')
    with _build_lock():
        if not force and _is_current(out_path, digest):
            return out_path
        build_dir = tempfile.mkdtemp(prefix=f'shim{os.getpid()}_', dir=_build_root())
        try:
            tmp = os.path.join(build_dir, SHIM_NAME)
            cuda_home = os.environ.get('CUDA_HOME', '/usr/local/cuda')
            incs = [os.path.join(os.path.dirname(torch.__file__), 'include'),
                    os.path.join(os.path.dirname(torch.__file__), 'include', 'torch', 'csrc', 'api',
                                 'include'),
                    os.path.join(cuda_home, 'include'), sysconfig.get_paths()['include'], INCLUDE]
            cmd = [gxx, '-O2', '-std=c++17', '-fPIC', '-shared', '-fvisibility=hidden',
                   f'-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}',
                   '-DTORCH_EXTENSION_NAME=_C', '-DTORCH_API_INCLUDE_EXTENSION_H', '-w']
            for i in incs:
                cmd += ['-isystem', i]
            cmd += [os.path.join(CSRC, SHIM_SOURCE), '-o', tmp]
            cmd += ['-L' + p for p in ce.library_paths()]
            cmd += ['-lc10', '-lc10_cuda', '-ltorch_cpu', '-ltorch_cuda', '-ltorch', '-ltorch_python',
                    '-ldl']
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                raise RuntimeError(f'g++ failed on {SHIM_SOURCE}:\n{res.stdout}\n{res.stderr}')
            _install(tmp, out_path, digest)
        finally:
            shutil.rmtree(build_dir, ignore_errors=True)
    return out_path


if __name__ == '__main__':
    if '--variant' in sys.argv:
        print(build_variant(int(sys.argv[sys.argv.index('--variant') + 1])))
        sys.exit(0)
    if '--shim' in sys.argv:
        print(build_shim(force='--force' in sys.argv))
        sys.exit(0)
    print(build(force='--force' in sys.argv, precise='--precise' in sys.argv,
                verbose='-v' in sys.argv, tune='--tune' in sys.argv))
