"""Builds l3c_pytorch_b200/libl3c_b200.so with nvcc for sm_100a (in-tree, no torch dependency).

    python -m l3c_pytorch_b200.build [-v]

nvcc cross-compiles without a GPU.  The shared library is git-ignored but travels to the GPU box
with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
SO = os.path.join(HERE, 'libl3c_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
SOURCES = ['capi.cu', 'range_coder.cu', 'dmll.cu', 'conv_ffma.cu', 'conv_tcgen05.cu', 'conv_f16.cu', 'conv_f16x2.cu', 'bicubic.cu', 'partition.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '--fmad=true',
              '-DL3C_BUILDING_DSO']


def _newest_source_mtime():
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    paths.append(os.path.join(os.path.dirname(HERE), 'include', 'l3c_b200.h'))
    return max(os.path.getmtime(p) for p in paths)


def build(verbose=False, force=False):
    if not force and os.path.isfile(SO) and os.path.getmtime(SO) >= _newest_source_mtime():
        return SO
    nvcc = NVCC
    objs = []
    build_dir = os.path.join(HERE, 'build')
    os.makedirs(build_dir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(build_dir, src.replace('.cu', '.o'))
        objs.append(obj)
        srcp = os.path.join(CSRC, src)
        deps = [srcp, os.path.join(CSRC, 'common.cuh'), os.path.join(CSRC, 'tc_ptx.cuh'),
                os.path.join(os.path.dirname(HERE), 'include', 'l3c_b200.h')]
        if not force and os.path.isfile(obj) and os.path.getmtime(obj) >= max(map(os.path.getmtime, deps)):
            continue
        cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', srcp, '-o', obj]
        if verbose:
            print(' '.join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for src, pr in procs:
        out = pr.communicate()[0].decode()
        if pr.returncode != 0:
            raise RuntimeError('nvcc failed for %s:\n%s' % (src, out))
        if verbose:
            print(out)
    cmd = [nvcc, '-shared', '-o', SO] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a',
                                                 '-lcudart_static', '-ldl', '-lrt', '-lpthread']
    subprocess.check_call(cmd)
    return SO


if __name__ == '__main__':
    print(build(verbose='-v' in sys.argv, force='-f' in sys.argv))
