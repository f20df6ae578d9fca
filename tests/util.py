"""Shared helpers for the parity tests."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def field_bits(a):
    """Bit pattern of every named field of a structured array (padding bytes excluded)."""
    a = np.ascontiguousarray(a)
    if a.dtype.names is None:
        return a.view(np.uint8).reshape(-1)
    return np.concatenate([np.ascontiguousarray(a[n]).view(np.uint8).reshape(-1) for n in a.dtype.names])


def same_bits(a, b):
    return np.array_equal(field_bits(a), field_bits(b))


def ulp_diff(a, b):
    """Max distance in units-in-the-last-place between two float32 arrays (0 == bit-identical)."""
    ia = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64)
    ib = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, np.int64(-2**31) - ia, ia)
    ib = np.where(ib < 0, np.int64(-2**31) - ib, ib)
    return int(np.abs(ia - ib).max()) if ia.size else 0


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))
