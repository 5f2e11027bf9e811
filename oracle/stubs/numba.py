"""Identity-decorator stand-in for numba (the JIT PEs are out of scope; oracle import aid only)."""
def _identity(*args, **kwargs):
    if len(args) == 1 and callable(args[0]) and not kwargs:
        return args[0]
    def deco(fn):
        return fn
    return deco
jit = njit = vectorize = _identity
prange = range
