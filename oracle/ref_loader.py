"""Import the UNMODIFIED reference (Python layer + its compiled Cython kernels) for parity work.

TEST INFRASTRUCTURE ONLY.  Two levels:

* ``load_kernels()``  -> (signal_functions, util, auto_interpretation): the reference's compiled
  Cython modules from oracle/_ref/ (built by oracle/build_ref.py).  These travel to the GPU box.
* ``load_python_layer()`` -> namespace with the reference's own ``Signal``, ``ProtocolAnalyzer``,
  ``Modulator``, ``Filter``, ``Spectrogram``, ``AutoInterpretation`` imported from
  /root/reference/src with a PyQt6 stub (no Qt in this image).  Only possible where
  /root/reference exists (the build container); used by tests/golden/make_golden.py to generate
  the committed fixtures.
"""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("URH_REFERENCE", "/root/reference")
REF_OUT = os.path.join(HERE, "_ref")
if not os.path.isdir(os.path.join(REF, "src", "urh")) and os.path.isdir(os.path.join(REF_OUT, "pyref", "src", "urh")):
    REF = os.path.join(REF_OUT, "pyref")   # the staged copy (oracle/build_ref.py: stage_python_layer) on the GPU box


def kernels_available() -> bool:
    try:
        load_kernels()
        return True
    except Exception:
        return False


def python_layer_available() -> bool:
    return os.path.isfile(os.path.join(REF, "src/urh/signalprocessing/Signal.py")) and kernels_available()


class _Sig:
    def __init__(self, *a, **k):
        pass

    def emit(self, *a, **k):
        pass

    def connect(self, *a, **k):
        pass

    def disconnect(self, *a, **k):
        pass


class _Meta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Dummy


class _Dummy(metaclass=_Meta):
    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Dummy()

    def __call__(self, *a, **k):
        return _Dummy()


class _QObject:
    def __init__(self, *a, **k):
        pass


class _QSettings:
    _store = {}
    class Format:
        IniFormat = 0
    class Scope:
        UserScope = 0
    IniFormat = 0
    UserScope = 0

    def __init__(self, *a, **k):
        pass

    def value(self, key, default=None, type=None):
        v = self._store.get(key, default)
        if type is not None and v is not None:
            try:
                return type(v)
            except Exception:
                return v
        return v

    def setValue(self, key, v):
        self._store[key] = v

    def fileName(self):
        return "/tmp/urh_oracle_settings.ini"

    def sync(self):
        pass

    def allKeys(self):
        return list(self._store)

    def contains(self, key):
        return key in self._store


class _QDir:
    @staticmethod
    def tempPath():
        import tempfile
        return tempfile.gettempdir()

    @staticmethod
    def homePath():
        return os.path.expanduser("~")


def _stub_module(name):
    m = types.ModuleType(name)
    m.__getattr__ = lambda attr: (_ for _ in ()).throw(AttributeError(attr)) if attr.startswith("__") else _Dummy
    return m


def _install_qt_stub():
    if "PyQt6" in sys.modules and getattr(sys.modules["PyQt6"], "_urh_oracle_stub", False):
        return
    pkg = types.ModuleType("PyQt6")
    pkg.__path__ = []
    pkg._urh_oracle_stub = True
    sys.modules["PyQt6"] = pkg
    for sub in ("QtCore", "QtGui", "QtWidgets", "QtTest", "uic", "QtSvg", "QtOpenGLWidgets"):
        m = _stub_module("PyQt6." + sub)
        sys.modules["PyQt6." + sub] = m
        setattr(pkg, sub, m)
    core = sys.modules["PyQt6.QtCore"]
    core.pyqtSignal = _Sig
    core.pyqtSlot = lambda *a, **k: (lambda f: f)
    core.QObject = _QObject
    core.QSettings = _QSettings
    core.QDir = _QDir


_kernels = None


def load_kernels():
    """The reference's compiled Cython modules, importable as urh.cythonext.* from oracle/_ref."""
    global _kernels
    if _kernels is not None:
        return _kernels
    if "urh" in sys.modules and not getattr(sys.modules["urh"], "__file__", "").startswith((REF_OUT, REF)):
        raise ImportError("a different 'urh' package is already imported")
    if not os.path.isdir(os.path.join(REF_OUT, "urh", "cythonext")):
        raise ImportError("oracle/_ref not built (run python oracle/build_ref.py)")
    if os.path.isdir(os.path.join(REF, "src", "urh")):
        # full reference python package first, compiled kernels grafted into urh.cythonext.__path__
        _install_qt_stub()
        if os.path.join(REF, "src") not in sys.path:
            sys.path.insert(0, os.path.join(REF, "src"))
        import urh.cythonext as ce
        p = os.path.join(REF_OUT, "urh", "cythonext")
        if p not in ce.__path__:
            ce.__path__.insert(0, p)
        if "urh.cythonext.path_creator" not in sys.modules:
            sys.modules["urh.cythonext.path_creator"] = types.ModuleType("urh.cythonext.path_creator")
    else:
        if REF_OUT not in sys.path:
            sys.path.insert(0, REF_OUT)
    sf = importlib.import_module("urh.cythonext.signal_functions")
    ut = importlib.import_module("urh.cythonext.util")
    ai = importlib.import_module("urh.cythonext.auto_interpretation")
    _kernels = (sf, ut, ai)
    return _kernels


def load_python_layer():
    """Reference Python DSP objects: from /root/reference in the build container, from the staged copy oracle/_ref/pyref
    on the GPU box."""
    load_kernels()
    if not os.path.isdir(os.path.join(REF, "src", "urh")):
        raise ImportError("reference python layer not present (neither /root/reference nor oracle/_ref/pyref)")
    ns = types.SimpleNamespace()
    from urh.signalprocessing.Signal import Signal
    from urh.signalprocessing.IQArray import IQArray
    from urh.signalprocessing.ProtocolAnalyzer import ProtocolAnalyzer
    from urh.signalprocessing.Modulator import Modulator
    from urh.signalprocessing.Filter import Filter, FilterType
    from urh.signalprocessing.Spectrogram import Spectrogram
    from urh.ainterpretation import AutoInterpretation, Wavelet
    ns.Signal, ns.IQArray, ns.ProtocolAnalyzer, ns.Modulator = Signal, IQArray, ProtocolAnalyzer, Modulator
    ns.Filter, ns.FilterType, ns.Spectrogram = Filter, FilterType, Spectrogram
    ns.AutoInterpretation, ns.Wavelet = AutoInterpretation, Wavelet
    ns.data_dir = os.path.join(REF, "tests", "data")
    return ns
