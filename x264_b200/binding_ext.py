"""argtypes for entry points beyond the pixel table; filled in as the translation units land."""
import ctypes as C


def bind(L):
    pass
