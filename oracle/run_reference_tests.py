#!/usr/bin/env python
"""Run the reference's OWN hot-path tests — unmodified files — through the reference's own Python layer
(Signal / ProtocolAnalyzer / Modulator / AutoInterpretation ...) with either

  --impl reference   the reference's compiled Cython kernels (oracle/_ref), or
  --impl b200        urh_b200's shims substituted for urh.cythonext.{signal_functions, util, auto_interpretation}
                     (the module substitution of INTEGRATION.md section 1): every public function the shim defines
                     replaces the reference's; what the shim does not define (CRC helpers, k_means ... outside the
                     IQ hot path) stays the reference's.

TEST INFRASTRUCTURE ONLY (lives under oracle/).  The reference comes from /root/reference in the build container and from the
staged copy oracle/_ref/pyref (oracle/build_ref.py) on the GPU box.

    python oracle/run_reference_tests.py --impl b200 [-- extra pytest args / test paths relative to the reference root]
"""
import argparse
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

DEFAULT_TESTS = ["tests/test_demodulations.py", "tests/test_modulator.py", "tests/test_iq_array.py", "tests/test_protocol_analyzer.py",
                 "tests/test_ringbuffer.py", "tests/auto_interpretation"]
MODULES = ["signal_functions", "util", "auto_interpretation"]


def substitute_b200():
    """composite urh.cythonext.* modules: the reference's compiled module, every public function of the urh_b200 shim on top"""
    import importlib

    import urh.cythonext as ce

    replaced = {}
    for short in MODULES:
        ref_mod = importlib.import_module("urh.cythonext." + short)
        shim = importlib.import_module("urh_b200.cythonext." + short)
        comp = types.ModuleType("urh.cythonext." + short)
        comp.__dict__.update({k: v for k, v in ref_mod.__dict__.items() if not k.startswith("__")})
        names = [k for k, v in shim.__dict__.items()
                 if not k.startswith("_") and callable(v) and getattr(v, "__module__", "") == shim.__name__]
        for k in names:
            setattr(comp, k, getattr(shim, k))
        comp.__urh_b200_substituted__ = sorted(names)
        sys.modules["urh.cythonext." + short] = comp
        setattr(ce, short, comp)
        replaced[short] = sorted(names)
    return replaced


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", choices=["reference", "b200"], default="b200")
    ap.add_argument("rest", nargs="*")
    args = ap.parse_args()
    sys.path.insert(0, HERE)
    import ref_loader

    ref_root = ref_loader.REF
    if not os.path.isdir(os.path.join(ref_root, "src", "urh")):
        print("REFERENCE_TESTS unavailable: no reference python layer (run python oracle/build_ref.py where /root/reference exists)")
        return 3
    # the reference root first (its `tests` package, its `src`), this repository last (urh_b200 only)
    sys.path.insert(0, ref_root)
    ref_loader.load_kernels()
    if ROOT not in sys.path:
        sys.path.append(ROOT)
    if args.impl == "b200":
        replaced = substitute_b200()
        for k, v in replaced.items():
            print("substituted urh.cythonext.%s: %s" % (k, ", ".join(v)))
        # make sure the substitution is what the reference's layer binds
        from urh.signalprocessing import Signal as S

        assert S.signal_functions.afp_demod.__module__.startswith("urh_b200"), "Signal.py did not bind the substituted module"
    import pytest

    tests = [a for a in args.rest if not a.startswith("-")] or DEFAULT_TESTS
    extra = [a for a in args.rest if a.startswith("-")]
    os.chdir(ref_root)
    rc = pytest.main(["-q", "-p", "no:cacheprovider", "--rootdir", ref_root, "-o", "python_files=test_*.py"] + extra + tests)
    print("REFERENCE_TESTS impl=%s rc=%d" % (args.impl, int(rc)))
    return int(rc)


if __name__ == "__main__":
    sys.exit(main())
