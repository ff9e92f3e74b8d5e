"""Build recipe for oracle/_ref/ -- TEST INFRASTRUCTURE ONLY.

Compiles the reference's own arithmetic coder, UNMODIFIED and from where it lies
(/root/reference/src/torchac/torchac_backend/torchac.cpp), into
oracle/_ref/torchac_backend_cpu.so.  The only thing added is `-DAT_CHECK=TORCH_CHECK`
(AT_CHECK was removed from modern libtorch).  No reference source is copied
into this repository; oracle/_ref/ is git-ignored but travels to the GPU box with gpurun.

The reference's own build system (src/torchac/setup.py) is not run.
"""
import os
import sys

REF_CPP = '/root/reference/src/torchac/torchac_backend/torchac.cpp'
HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, '_ref')


def ref_so_path():
    return os.path.join(OUT_DIR, 'torchac_backend_cpu.so')


def build(verbose=False):
    """Returns the path of the built module, or None if the reference mount is absent."""
    if os.path.isfile(ref_so_path()) and not os.path.isfile(REF_CPP):
        return ref_so_path()          # GPU box: prebuilt file travelled with the snapshot
    if not os.path.isfile(REF_CPP):
        return None
    if os.path.isfile(ref_so_path()) and \
            os.path.getmtime(ref_so_path()) >= os.path.getmtime(REF_CPP):
        return ref_so_path()
    os.makedirs(OUT_DIR, exist_ok=True)
    from torch.utils import cpp_extension
    cpp_extension.load(
        name='torchac_backend_cpu', sources=[REF_CPP], build_directory=OUT_DIR,
        extra_cflags=['-O3', '-DAT_CHECK=TORCH_CHECK'], verbose=verbose, is_python_module=False)
    return ref_so_path()


def load():
    """Import the compiled reference backend (None if it is not available)."""
    p = build()
    if p is None or not os.path.isfile(p):
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location('torchac_backend_cpu', p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == '__main__':
    print(build(verbose='-v' in sys.argv))
