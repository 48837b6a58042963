"""Drop-in modules for ``urh.cythonext.{signal_functions, util, auto_interpretation}`` on the IQ hot path.

Same function names, positional/keyword arguments, return types and exceptions as the reference's
Cython modules — the work runs in hand-written CUDA (liburh_b200.so) instead of the Cython/OpenMP loops.
Host numpy arrays in -> numpy arrays out; ``urh_b200.device.DeviceArray`` in -> DeviceArray out (the data
stays resident in HBM between calls).
"""
