"""Recipe: compile the REFERENCE's own CPU backend into oracle/_ref/ (test infrastructure only).

TEST INFRASTRUCTURE — only tests/, __graft_entry__.smoke()/build() and bench.py's
cpu_baseline / --impl reference legs may use what this produces.

What it does
------------
Compiles the six translation units that the reference's setup.py lists for the
``sige.cpu`` extension (reference setup.py:153-163: sige/cpu/{gather,scatter,
scatter_gather,common_cpu,pybind_cpu}.cpp + sige/common.cpp) **from where they lie
under /root/reference** with the flags of reference setup.py:148-150
(``-g -O3 -fopenmp``) into ``oracle/_ref/sige_ref_cpu.so``.  No reference source
is copied into this repository; only the built shared object lands in
``oracle/_ref/`` (git-ignored, but shipped to the GPU box by gpurun).

The module exports the reference pybind entry points (reference
sige/cpu/pybind_cpu.cpp:5-12): gather, scatter, scatter_with_block_residual,
scatter_gather, get_scatter_map.

The default ``/opt/gcc/bin/g++`` wrapper in this image cannot find libgomp.spec, so
the recipe pins CXX=/usr/bin/g++.
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("SIGE_REFERENCE_ROOT", "/root/reference")
OUT_DIR = os.path.join(HERE, "_ref")
MODULE_NAME = "sige_ref_cpu"

_SOURCES = [
    "sige/cpu/gather.cpp",
    "sige/cpu/scatter.cpp",
    "sige/cpu/scatter_gather.cpp",
    "sige/cpu/common_cpu.cpp",
    "sige/cpu/pybind_cpu.cpp",
    "sige/common.cpp",
]


def ref_available() -> bool:
    return all(os.path.isfile(os.path.join(REF_ROOT, s)) for s in _SOURCES)


def built_path() -> str:
    return os.path.join(OUT_DIR, MODULE_NAME + ".so")


def build(force: bool = False, verbose: bool = False) -> str | None:
    """Build oracle/_ref/sige_ref_cpu.so if the reference tree is present.

    Returns the path of the shared object, or None when /root/reference is absent
    (the GPU box): then the prebuilt file, if it travelled, is used as is.
    """
    so = built_path()
    if os.path.isfile(so) and not force:
        return so
    if not ref_available():
        return so if os.path.isfile(so) else None
    os.makedirs(OUT_DIR, exist_ok=True)
    os.environ["CXX"] = "/usr/bin/g++"
    os.environ["CC"] = "/usr/bin/gcc"
    from torch.utils.cpp_extension import load

    build_dir = os.path.join(OUT_DIR, "_build")
    os.makedirs(build_dir, exist_ok=True)
    load(
        name=MODULE_NAME,
        sources=[os.path.join(REF_ROOT, s) for s in _SOURCES],
        extra_cflags=["-g", "-O3", "-fopenmp"],
        extra_ldflags=["-fopenmp"],
        build_directory=build_dir,
        is_python_module=False,
        verbose=verbose,
    )
    shutil.copy2(os.path.join(build_dir, MODULE_NAME + ".so"), so)
    shutil.rmtree(build_dir, ignore_errors=True)
    return so


def load_ref():
    """Import the compiled reference CPU backend as a python module (or None)."""
    so = build()
    if so is None or not os.path.isfile(so):
        return None
    import importlib.util

    import torch  # noqa: F401  (libtorch must be loaded before the extension)

    spec = importlib.util.spec_from_file_location(MODULE_NAME, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose=True)
    print("reference CPU backend:", p)
