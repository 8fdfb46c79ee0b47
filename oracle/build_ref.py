"""TEST INFRASTRUCTURE — compiles the reference's ONLY native file on this path, sam2/csrc/connected_components.cu
(the extension behind sam2._C.get_connected_componnets, misc.py:48-61), from the sources WHERE THEY LIE under
/root/reference into oracle/_ref/ (git-ignored; travels to the GPU box with the repo snapshot like every built .so).

The file is a PyTorch C++/CUDA extension (ATen + pybind11), so the recipe is torch.utils.cpp_extension with an explicit
sm_100a target — nvcc cross-compiles here without a GPU; the reference's own build system (setup.py) is not run and no
reference source is copied.  On the GPU box tests/test_cc_reference_gpu.py loads the module and pins this repo's
ds2_connected_components / ds2_fill_holes (and oracle/cc_oracle.c) against the reference kernel itself.

    python -m oracle.build_ref          # -> oracle/_ref/sam2_ref_C.so
"""
import glob
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_ROOT = os.environ.get("DS2_REFERENCE_ROOT", "/root/reference")
SRC = os.path.join(REF_ROOT, "sam2", "csrc", "connected_components.cu")
OUT_DIR = os.path.join(ROOT, "oracle", "_ref")
NAME = "sam2_ref_C"


def built_path():
    hits = glob.glob(os.path.join(OUT_DIR, NAME + "*.so"))
    return hits[0] if hits else None


def build(force=False):
    """Returns the path of the built module, or None when the reference sources are absent (GPU box)."""
    have = built_path()
    if have and not force:
        return have
    if not os.path.exists(SRC):
        return None
    os.makedirs(OUT_DIR, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils import cpp_extension
    work = os.path.join(OUT_DIR, "_work")
    os.makedirs(work, exist_ok=True)
    cpp_extension.load(name=NAME, sources=[SRC], build_directory=work, is_python_module=False, verbose=False,
                       extra_cuda_cflags=["-gencode=arch=compute_100a,code=sm_100a", "-O2"], with_cuda=True)
    so = os.path.join(work, NAME + ".so")
    shutil.copy2(so, os.path.join(OUT_DIR, NAME + ".so"))
    shutil.rmtree(work, ignore_errors=True)
    return built_path()


def load():
    """Imports the built extension (needs torch; the module's function needs a CUDA device to run)."""
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded before the extension)
    path = built_path()
    if path is None:
        raise FileNotFoundError("oracle/_ref/sam2_ref_C.so missing: run `python -m oracle.build_ref` in the build container")
    spec = importlib.util.spec_from_file_location(NAME, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build(force="--force" in sys.argv)
    print(p or f"reference source not found at {SRC}")
