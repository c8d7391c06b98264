"""Build recipe for ``oracle/_ref`` -- the UNMODIFIED reference Chamfer CUDA extension.

TEST INFRASTRUCTURE ONLY.  Nothing under ``softpool_b200/`` may import this.

Compiles the reference's own sources *where they lie* under ``/root/reference``
(``distance/chamfer/chamfer.cu`` + ``chamfer_cuda.cpp``) for sm_100 with
``torch.utils.cpp_extension`` into ``oracle/_ref/`` (git-ignored, NOT
gpurun-ignored, so the built ``.so`` travels to the GPU box).  No reference
source is copied into this repository.

The reference kernels can only *run* on a GPU, so the resulting module is used by
``tests/test_chamfer_gpu.py`` on the B200 box to pin the C restatement in
``oracle/chamfer_oracle.c`` (and, through it, our kernels) against the reference
itself, and by ``bench.py --ref-gpu`` as "the kernel to beat".

``stage_softpool()`` additionally stages the reference's ``softpool.py`` (one file, copied verbatim, NOT into
git: ``oracle/_ref/`` is git-ignored) as ``oracle/_ref/softpool_ref.py`` so that ``bench.py --impl reference`` and
its ``ref_gpu`` leg can run the reference module ITSELF on the GPU box, where /root/reference does not exist.

Usage:  python oracle/build_ref.py            (no-op when /root/reference is absent)
"""
import glob
import os
import sys

REF = "/root/reference/distance/chamfer"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
NAME = "softpool_ref_chamfer"


def built_path():
    hits = glob.glob(os.path.join(OUT, NAME + "*.so"))
    return hits[0] if hits else None


def build(verbose=False):
    if built_path():
        return built_path()
    if not os.path.isdir(REF):
        return None
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils.cpp_extension import load
    load(name=NAME,
         sources=[os.path.join(REF, "chamfer_cuda.cpp"), os.path.join(REF, "chamfer.cu")],
         build_directory=OUT, verbose=verbose, is_python_module=False,
         extra_cuda_cflags=["-lineinfo"])
    return built_path()


REF_SOFTPOOL = "/root/reference/softpool.py"
STAGED_SOFTPOOL = os.path.join(OUT, "softpool_ref.py")


def stage_softpool():
    """Copy the reference softpool.py verbatim into oracle/_ref (git-ignored); returns the staged path or None."""
    if os.path.exists(STAGED_SOFTPOOL):
        return STAGED_SOFTPOOL
    if not os.path.exists(REF_SOFTPOOL):
        return None
    os.makedirs(OUT, exist_ok=True)
    import shutil
    shutil.copyfile(REF_SOFTPOOL, STAGED_SOFTPOOL)
    return STAGED_SOFTPOOL


def load_ref_softpool(cpu):
    """Import the staged reference softpool.py.  cpu=True neutralises the hard-coded `.cuda()` calls (softpool.py:24,...)
    to the identity BEFORE the import -- no line of the reference is changed; cpu=False imports it as it is."""
    p = STAGED_SOFTPOOL if os.path.exists(STAGED_SOFTPOOL) else (REF_SOFTPOOL if os.path.exists(REF_SOFTPOOL) else None)
    if p is None:
        return None
    import importlib.util
    import torch
    import torch.nn as nn
    if cpu:
        torch.Tensor.cuda = lambda self, *a, **k: self
        nn.Module.cuda = lambda self, *a, **k: self
    spec = importlib.util.spec_from_file_location("softpool_ref_cpu" if cpu else "softpool_ref_cuda", p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_ref():
    """Import the built reference extension (GPU box or here); None if it was never built."""
    p = built_path()
    if p is None:
        return None
    import importlib.util
    import torch  # noqa: F401  (registers libtorch symbols the extension links against)
    spec = importlib.util.spec_from_file_location(NAME, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    stage_softpool()
    p = build(verbose="-v" in sys.argv)
    print("oracle/_ref:", p if p else "unavailable (no /root/reference)")
