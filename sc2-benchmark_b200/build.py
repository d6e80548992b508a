"""Builds libsc2b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python sc2-benchmark_b200/build.py [--force]

The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_DIR = os.path.join(HERE, 'lib')
LIB_PATH = os.path.join(LIB_DIR, 'libsc2b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden', '--use_fast_math=false']
NVCC_FLAGS.remove('--use_fast_math=false')  # never fast-math: bit-exactness matters on this path
NVCC_FLAGS += os.environ.get('SC2_NVCC_DEFINES', '').split()  # diagnostics builds, e.g. -DSC2_HANG_DEBUG


HOSTBYTES_SRC = os.path.join(CSRC, 'hostbytes.c')


def hostbytes_path():
    import sysconfig
    return os.path.join(LIB_DIR, '_sc2_hostbytes' + sysconfig.get_config_var('EXT_SUFFIX'))


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def build_hostbytes():
    """The CPython helper that splits / gathers the contract's list[bytes] with the GIL released (csrc/hostbytes.c)."""
    import sysconfig
    out = hostbytes_path()
    cc = os.environ.get('CC', 'gcc')
    subprocess.check_call([cc, '-O2', '-shared', '-fPIC', '-I', sysconfig.get_paths()['include'], HOSTBYTES_SRC, '-o', out])
    return out


def _fingerprint():
    h = hashlib.sha256()
    for path in _sources() + sorted(glob.glob(os.path.join(CSRC, '*.cuh'))) + \
            [HOSTBYTES_SRC, os.path.join(os.path.dirname(HERE), 'include', 'sc2b200.h'), os.path.abspath(__file__)]:
        h.update(path.encode())
        with open(path, 'rb') as f:
            h.update(f.read())
    return h.hexdigest()


def build_native(force=False, verbose=False):
    os.makedirs(LIB_DIR, exist_ok=True)
    stamp = os.path.join(LIB_DIR, 'build.stamp')
    fp = _fingerprint()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(hostbytes_path()) and os.path.exists(stamp) \
            and open(stamp).read().strip() == fp:
        return LIB_PATH
    if not os.path.exists(NVCC):
        raise RuntimeError('nvcc not found at %s; libsc2b200.so cannot be built' % NVCC)
    objs = []
    procs = []
    for src in _sources():
        obj = os.path.join(LIB_DIR, os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        cmd = [NVCC] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for cmd, proc in procs:
        out, _ = proc.communicate()
        if verbose or proc.returncode != 0:
            sys.stderr.write(out.decode())
        if proc.returncode != 0:
            raise RuntimeError('nvcc failed: ' + ' '.join(cmd))
    link = [NVCC, '-shared', '-o', LIB_PATH] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-Xcompiler', '-fPIC']
    subprocess.check_call(link)
    build_hostbytes()
    with open(stamp, 'w') as f:
        f.write(fp)
    return LIB_PATH


if __name__ == '__main__':
    print(build_native(force='--force' in sys.argv, verbose='-v' in sys.argv))
