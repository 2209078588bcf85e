"""ctypes wrapper of oracle/b2d_oracle.c. TEST INFRASTRUCTURE ONLY."""
import ctypes

import numpy as np

from .build import build_oracle

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_oracle())
        _lib.oracle_bf16_round.restype = ctypes.c_float
        _lib.oracle_bf16_round.argtypes = [ctypes.c_float]
        _lib.oracle_wire_bf16.restype = ctypes.c_float
        _lib.oracle_wire_bf16.argtypes = [ctypes.c_float, ctypes.c_float]
    return _lib


def _ptrs(arrs):
    arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in arrs]
    P = (ctypes.POINTER(ctypes.c_float) * len(arrs))(*[a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)) for a in arrs])
    return arrs, P


def allreduce(per_rank, wire="bf16", scale=None):
    world = len(per_rank)
    scale = np.float32(1.0) / np.float32(world) if scale is None else np.float32(scale)
    arrs, P = _ptrs(per_rank)
    out = np.empty_like(arrs[0])
    fn = lib().oracle_allreduce_bf16 if wire == "bf16" else lib().oracle_allreduce_fp32
    fn(P, ctypes.c_int(world), ctypes.c_size_t(out.size), ctypes.c_float(float(scale)),
       out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    return out


def adam(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, adamw=False):
    fp = ctypes.POINTER(ctypes.c_float)
    for a in (p, g, m, v):
        assert a.dtype == np.float32 and a.flags.c_contiguous
    lib().oracle_adam(p.ctypes.data_as(fp), g.ctypes.data_as(fp), m.ctypes.data_as(fp), v.ctypes.data_as(fp),
                      ctypes.c_size_t(p.size), ctypes.c_int(step), ctypes.c_double(lr), ctypes.c_double(beta1),
                      ctypes.c_double(beta2), ctypes.c_double(eps), ctypes.c_double(weight_decay),
                      ctypes.c_int(int(adamw)))
    return p, m, v


def bucket_assignment(numels, limits, elem_size=4, reverse=True):
    n = len(numels)
    a = (ctypes.c_int64 * n)(*numels)
    lim = (ctypes.c_int64 * len(limits))(*limits)
    out = (ctypes.c_int * n)()
    nb = lib().oracle_bucket_assignment(a, n, elem_size, lim, len(limits), out)
    buckets = [[] for _ in range(nb)]
    for i in range(n):
        buckets[out[i]].append(i)
    return list(reversed(buckets)) if reverse else buckets


def partition_fairscale(numels, world):
    n = len(numels)
    a = (ctypes.c_int64 * n)(*numels)
    out = (ctypes.c_int * n)()
    lib().oracle_partition_fairscale(a, n, world, out)
    return list(out)
