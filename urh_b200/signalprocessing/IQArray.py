"""``IQArray`` — (n,2) C-contiguous sample container (reference: src/urh/signalprocessing/IQArray.py:11-319).

Same constructor, properties, dtype-conversion rules (pinned by the reference's tests/test_iq_array.py) and
file formats.  `magnitudes` runs on the GPU (util.get_magnitudes); `device()` returns / caches the capture in HBM
so that Signal, Filter and the demodulators do not re-upload it.
"""
import os
import tarfile
import tempfile
import wave

import numpy as np

from ..cythonext.util import get_magnitudes

_INT_TYPES = (np.uint8, np.int8, np.uint16, np.int16)


class IQArray(object):
    def __init__(self, data: np.ndarray, dtype=None, n=None, skip_conversion=False, _owned=False):
        if data is None:
            self.__data = np.zeros((n, 2), dtype, order="C")
        elif skip_conversion:
            self.__data = data
        else:
            self.__data = self.convert_array_to_iq(data)
        assert self.__data.dtype not in (np.complex64, np.complex128)
        self._device = None
        # the caller may keep (and later write) the array it passed in unless we made our own copy
        self._aliased = data is not None and not _owned   # _owned: the caller hands the array over and keeps no reference

    # -- numpy-like access -------------------------------------------------------------------------------------
    # The HBM copy (`device()`) must never go stale.  Every accessor that hands out a WRITABLE numpy view of the samples
    # (`data`, `real`, `imag`, `iq[...]`, `convert_to` of the same dtype — the reference lets callers write through them,
    # e.g. ``iq.data[a:b] = 0``) marks the object as aliased: the view may be written at any later time, so from then on
    # `device()` uploads afresh on every call instead of trusting a cached copy.  Readers inside the package use
    # `_peek()`, a read-only view, and keep the cache.
    def _alias(self):
        self._device = None
        self._aliased = True

    def __getitem__(self, item):
        self._alias()
        return self.__data[item]

    def _peek(self, item=None):
        view = self.__data.view() if item is None else self.__data[item]
        if isinstance(view, np.ndarray):
            view.flags.writeable = False
        return view

    def __setitem__(self, key, value):
        self._device = None
        if isinstance(value, (int, float)):
            self.__data[key] = value
            return
        if isinstance(value, IQArray):
            value = value.data
        if value.dtype == np.complex64 or value.dtype == np.complex128:
            self.real[key] = value.real
            self.imag[key] = value.imag
        elif value.ndim == 2:
            self.__data[key] = value
        else:
            self.__data[key] = value.reshape((-1, 2), order="C")

    def __len__(self):
        return len(self.__data)

    def __eq__(self, other):
        return np.array_equal(self._peek(), other._peek() if isinstance(other, IQArray) else other.data)

    @property
    def num_samples(self):
        return self.__data.shape[0]

    @property
    def minimum(self):
        return self.min_max_for_dtype(self.__data.dtype)[0]

    @property
    def maximum(self):
        return self.min_max_for_dtype(self.__data.dtype)[1]

    @property
    def data(self):
        self._alias()
        return self.__data

    @property
    def real(self):
        self._alias()
        return self.__data[:, 0]

    @real.setter
    def real(self, value):
        self._device = None
        self.__data[:, 0] = value

    @property
    def imag(self):
        self._alias()
        return self.__data[:, 1]

    @imag.setter
    def imag(self, value):
        self._device = None
        self.__data[:, 1] = value

    @property
    def dtype(self):
        return self.__data.dtype

    # -- GPU residency ---------------------------------------------------------------------------------------------
    def device(self):
        """the capture as a DeviceArray (uploaded once, invalidated by in-place edits through this object)"""
        from ..device import to_device

        if self._aliased:
            return to_device(np.ascontiguousarray(self.__data))   # a writable view is out there: never cache
        if self._device is None or len(self._device) != len(self.__data):
            self._device = to_device(np.ascontiguousarray(self.__data))
        return self._device

    def own(self):
        """take a private copy of the samples: no outside view can reach them any more, so `device()` may cache again"""
        self.__data = np.array(self.__data, order="C")
        self._device = None
        self._aliased = False
        return self

    @property
    def magnitudes(self):
        return get_magnitudes(np.ascontiguousarray(self.__data))

    @property
    def magnitudes_normalized(self):
        return self.magnitudes / np.sqrt(self.maximum ** 2.0 + self.minimum ** 2.0)

    def as_complex64(self):
        return self.convert_to(np.float32).flatten(order="C").view(np.complex64)

    def to_bytes(self):
        return self.__data.tobytes()

    def subarray(self, start=None, stop=None, step=None):
        return IQArray(np.array(self._peek(slice(start, stop, step)), order="C"), _owned=True)

    def insert_subarray(self, pos, subarray: np.ndarray):
        self._device = None
        if subarray.ndim == 1:
            if subarray.dtype == np.complex64:
                subarray = subarray.view(np.float32)
            elif subarray.dtype == np.complex128:
                subarray = subarray.view(np.float64)
            subarray = subarray.reshape((-1, 2), order="C")
        self.__data = np.insert(self.__data, pos, subarray, axis=0)

    def apply_mask(self, mask: np.ndarray):
        self._device = None
        self.__data = self.__data[mask]

    # -- dtype conversion (IQArray.py:129-203) -------------------------------------------------------------------------
    def convert_to(self, target_dtype) -> np.ndarray:
        """IQArray.py:127-200.  The element-wise conversion runs on the GPU (convert.cu); the same object is returned when
        the dtype already matches."""
        tgt = np.dtype(target_dtype)
        if tgt == self.__data.dtype:
            self._alias()
            return self.__data
        if tgt not in [np.dtype(t) for t in _INT_TYPES + (np.float32,)]:
            raise ValueError("Data type {} not supported".format(target_dtype))
        return self.convert_to_device(tgt).get()

    def convert_to_device(self, target_dtype):
        """the converted capture as a DeviceArray (no download)"""
        import ctypes as C

        from .. import _lib
        from ..device import DeviceArray

        tgt = np.dtype(target_dtype)
        src = self.device()
        if tgt == src.dtype:
            return src
        out = DeviceArray(src.ctx, self.__data.shape, tgt)
        src.ctx.check(src.ctx.lib.urh_convert_iq(src.ctx.handle, C.c_void_p(src.ptr), _lib.dtype_code(src.dtype), C.c_void_p(out.ptr),
                                                  _lib.dtype_code(tgt), int(self.__data.size)))
        return out

    # -- files (IQArray.py:115-127, 205-227, 263-275) -------------------------------------------------------------------
    _EXT = {
        (".complex16u", ".cu8"): np.uint8,
        (".complex16s", ".cs8"): np.int8,
        (".complex32u", ".cu16"): np.uint16,
        (".complex32s", ".cs16"): np.int16,
    }

    @classmethod
    def _dtype_for_filename(cls, filename: str):
        for exts, dt in cls._EXT.items():
            if filename.endswith(exts):
                return dt
        return np.float32

    def tofile(self, filename: str):
        self.convert_to(self._dtype_for_filename(filename)).tofile(filename)

    @staticmethod
    def from_file(filename: str):
        dt = IQArray._dtype_for_filename(filename)
        arr = IQArray(data=np.fromfile(filename, dtype=dt), _owned=True)
        if dt == np.uint8:
            return IQArray(arr.convert_to(np.int8), _owned=True)      # unsigned captures are handled as signed
        if dt == np.uint16:
            return IQArray(arr.convert_to(np.int16), _owned=True)
        return arr

    @staticmethod
    def convert_array_to_iq(arr: np.ndarray) -> np.ndarray:
        if arr.ndim == 1:
            if arr.dtype == np.complex64:
                arr = arr.view(np.float32)
            elif arr.dtype == np.complex128:
                arr = arr.view(np.float64)
            if len(arr) % 2:
                arr = arr[:-1]  # drop a trailing half sample
            return arr.reshape((-1, 2), order="C")
        if arr.ndim == 2:
            return arr
        raise ValueError("Too many dimensions")

    @staticmethod
    def min_max_for_dtype(dtype) -> tuple:
        if dtype in (np.float32, np.float64, np.complex64, np.complex128):
            return -1, 1
        return np.iinfo(dtype).min, np.iinfo(dtype).max

    @staticmethod
    def concatenate(*args):
        return IQArray(data=np.concatenate([a._peek() if isinstance(a, IQArray) else a for a in args[0]]), _owned=True)

    def save_compressed(self, filename):
        with tarfile.open(filename, "w:bz2") as tar_write:
            tmp_name = tempfile.mkstemp()[1]
            self.tofile(tmp_name)
            tar_write.add(tmp_name)
        os.remove(tmp_name)

    def export_to_wav(self, filename, num_channels, sample_rate):
        f = wave.open(filename, "w")
        f.setnchannels(num_channels)
        f.setsampwidth(2)
        f.setframerate(sample_rate)
        f.writeframes(self.convert_to(np.int16))
        f.close()
