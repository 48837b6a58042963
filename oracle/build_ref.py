"""Build the reference's own Cython DSP kernels into oracle/_ref/ (TEST INFRASTRUCTURE ONLY).

Compiles, unmodified and where they lie, the three hot-path .pyx files of the reference
(/root/reference/src/urh/cythonext/{signal_functions,util,auto_interpretation}.pyx) with the
reference's own compiler directives (setup.py:128-131, dev/native/ExtensionHelper.py:14-20:
language_level=3, cdivision, wraparound=False, boundscheck=False, initializedcheck=False),
language=c++ and OpenMP (setup.py:104-114).  Generated C++ and the .so files go ONLY to
oracle/_ref/ (git-ignored, but shipped to the GPU box by gpurun).  No reference source is
copied into the repository.

Usage:  python oracle/build_ref.py            (no-op if /root/reference is absent)
"""
import os
import shutil
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("URH_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
PYX = ["signal_functions", "util", "auto_interpretation"]


def ref_available() -> bool:
    return os.path.isfile(os.path.join(REF, "src/urh/cythonext/signal_functions.pyx"))


def built() -> bool:
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    return all(os.path.isfile(os.path.join(OUT, "urh", "cythonext", m + suffix)) for m in PYX)


def build(force: bool = False) -> bool:
    if not ref_available():
        return built()
    if built() and not force:
        return True
    # The image's default CC wrapper cannot link -fopenmp (libgomp.spec missing): use /usr/bin.
    os.environ["CC"] = "/usr/bin/gcc"
    os.environ["CXX"] = "/usr/bin/g++"
    os.environ["LDSHARED"] = "/usr/bin/g++ -shared"
    import numpy as np
    from Cython.Build import cythonize
    from setuptools import Extension
    from setuptools.dist import Distribution
    from setuptools.command.build_ext import build_ext

    src_dir = os.path.join(REF, "src")
    pkg_dir = os.path.join(OUT, "urh", "cythonext")
    os.makedirs(pkg_dir, exist_ok=True)
    for d in (os.path.join(OUT, "urh"), pkg_dir):
        init = os.path.join(d, "__init__.py")
        if not os.path.exists(init):
            open(init, "w").close()
    exts = [
        Extension(
            "urh.cythonext." + m,
            [os.path.join(src_dir, "urh", "cythonext", m + ".pyx")],
            extra_compile_args=["-fopenmp", "-O2", "-Wno-cpp", "-w"],
            extra_link_args=["-fopenmp"],
            include_dirs=[np.get_include()],
            language="c++",
        )
        for m in PYX
    ]
    cwd = os.getcwd()
    os.chdir(src_dir)  # so that "urh.cythonext.util" cimports resolve; nothing is written here
    try:
        exts = cythonize(
            exts,
            build_dir=os.path.join(OUT, "_gen"),
            include_path=[src_dir],
            compiler_directives=dict(
                language_level=3, cdivision=True, wraparound=False, boundscheck=False, initializedcheck=False
            ),
            quiet=True,
        )
        dist = Distribution({"ext_modules": exts})
        cmd = build_ext(dist)
        cmd.build_lib = OUT
        cmd.build_temp = os.path.join(OUT, "_tmp")
        cmd.inplace = 0
        cmd.ensure_finalized()
        cmd.run()
    finally:
        os.chdir(cwd)
    shutil.rmtree(os.path.join(OUT, "_tmp"), ignore_errors=True)
    shutil.rmtree(os.path.join(OUT, "_gen"), ignore_errors=True)
    return built()


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref built:", ok)
    sys.exit(0 if ok or not ref_available() else 1)
