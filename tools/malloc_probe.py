#!/usr/bin/env python
"""How long does the first training wave's device allocation take as 1536 cudaMalloc calls (192 regions x 8 blocks, as
sml_train_begin issues them) against one arena of the same total?  Prints one JSON line."""
import ctypes as C
import json
import time

rt = C.CDLL("/usr/local/cuda/lib64/libcudart.so")
rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
rt.cudaFree.argtypes = [C.c_void_p]
rt.cudaFree(None)                      # context
ld, ks, n = 6032, 1024, 5760
sizes = [ld * ld * 8, ld * ks * 8, n * 8, n * 8, 47 * 2 * 128 * 128 * 8, ld * 8, 4, 136 * 4]


def many(nreg):
    ptrs = []
    t0 = time.perf_counter()
    for _ in range(nreg):
        for s in sizes:
            p = C.c_void_p()
            assert rt.cudaMalloc(C.byref(p), s) == 0
            ptrs.append(p)
    t1 = time.perf_counter()
    for p in ptrs:
        rt.cudaFree(p)
    return t1 - t0, time.perf_counter() - t1


def arena(nreg):
    total = sum((s + 255) // 256 * 256 for s in sizes) * nreg
    p = C.c_void_p()
    t0 = time.perf_counter()
    assert rt.cudaMalloc(C.byref(p), total) == 0
    t1 = time.perf_counter()
    rt.cudaFree(p)
    return t1 - t0, time.perf_counter() - t1, total


def memset_first_touch(nreg):
    """allocate like a wave, then zero every Gram block (what sml_train_begin enqueues) and wait"""
    rt.cudaMemsetAsync.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p]
    ptrs = []
    for _ in range(nreg):
        p = C.c_void_p()
        assert rt.cudaMalloc(C.byref(p), sizes[0]) == 0
        ptrs.append(p)
    res = []
    for _ in range(2):                      # first touch, then again
        t0 = time.perf_counter()
        for p in ptrs:
            rt.cudaMemsetAsync(p, 0, sizes[0], None)
        rt.cudaDeviceSynchronize()
        res.append(time.perf_counter() - t0)
    for p in ptrs:
        rt.cudaFree(p)
    return res


out = {"memset_192_grams_s": memset_first_touch(192)}
for nreg in (48, 192):
    a, fa, total = arena(nreg)
    m, fm = many(nreg)
    a2, _, _ = arena(nreg)
    out[str(nreg)] = {"bytes": total, "many_malloc_s": m, "many_free_s": fm, "arena_malloc_s": a, "arena_free_s": fa, "arena_again_s": a2}
print(json.dumps(out))
