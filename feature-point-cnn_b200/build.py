"""Build libspb200.so (C ABI + CUDA kernels, sm_100a only) in-tree with nvcc.

    python feature-point-cnn_b200/build.py [--force]

Objects go to feature-point-cnn_b200/build/, the library to feature-point-cnn_b200/libspb200.so.
nvcc cross-compiles without a GPU; the .so travels to the GPU box with the repository snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(HERE, 'build')
LIB = os.path.join(HERE, 'libspb200.so')
INCLUDE = os.path.join(os.path.dirname(HERE), 'include')
SOURCES = ['ckpt_reader.cpp', 'conv_simt.cu', 'conv_tc.cu', 'block_tc.cu', 'halo_tc.cu', 'stem_tc.cu', 'stem_planes.cu', 'nms.cu', 'match.cu', 'match_tc.cu', 'homography.cu', 'postproc.cu', 'preproc.cu', 'engine.cu', 'capi.cu']
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-Xcompiler', '-fPIC',
         '-Xcompiler', '-fvisibility=hidden', '-I', INCLUDE, '-I', CSRC]


def _newest_header():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.h', '.cuh'))]
    hs.append(os.path.join(INCLUDE, 'spb200.h'))
    return max(os.path.getmtime(h) for h in hs)


def _compile(src, force, verbose):
    obj = os.path.join(OBJ, os.path.splitext(src)[0] + '.o')
    path = os.path.join(CSRC, src)
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(path), _newest_header()):
        return obj
    cmd = [NVCC] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', path, '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError('nvcc failed on ' + src)
    return obj


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    with ThreadPoolExecutor(max_workers=6) as ex:
        objs = list(ex.map(lambda s: _compile(s, force, verbose), SOURCES))
    if force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        subprocess.check_call([NVCC, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a'])
    build_cpp_demo(force)
    return LIB


def build_cpp_demo(force=False):
    """The C++ facade (cpp/superpoint.h, header-only over the C ABI) compiled into its headless demo."""
    cpp = os.path.join(HERE, 'cpp')
    exe, src, hdr = os.path.join(cpp, 'demo'), os.path.join(cpp, 'demo.cc'), os.path.join(cpp, 'superpoint.h')
    if not force and os.path.exists(exe) and os.path.getmtime(exe) >= max(os.path.getmtime(src), os.path.getmtime(hdr), os.path.getmtime(LIB)):
        return exe
    subprocess.check_call(['g++', '-std=c++17', '-O2', '-Wall', '-Wextra', '-Werror', '-I', INCLUDE, '-I', cpp, src, '-L', HERE,
                           '-lspb200', '-Wl,-rpath,$ORIGIN/..', '-o', exe])
    return exe


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
