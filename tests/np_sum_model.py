"""Pure-Python model of numpy's float32 pairwise summation (np.sum of a contiguous 1-D float32
array) -- the order of additions csrc/kab_segstats.cuh reproduces on the device for align()'s
per-segment sums (align.py:161-162).  Test infrastructure: the CPU suite pins this model against
np.sum itself, the GPU suite compares the kernel with np.sum directly."""
import numpy as np

f32 = np.float32


def _block(a):
    n = len(a)
    if n < 8:
        r = f32(0.0)
        for x in a:
            r = f32(r + x)
        return r
    r = [f32(a[j]) for j in range(8)]
    i = 8
    while i + 8 <= n:
        for j in range(8):
            r[j] = f32(r[j] + a[i + j])
        i += 8
    res = f32(f32(f32(r[0] + r[1]) + f32(r[2] + r[3])) + f32(f32(r[4] + r[5]) + f32(r[6] + r[7])))
    while i < n:
        res = f32(res + a[i])
        i += 1
    return res


def pairwise_sum(a):
    a = np.asarray(a, dtype=np.float32)
    n = len(a)
    if n <= 128:
        return _block(a)
    n2 = n // 2
    n2 -= n2 % 8
    return f32(pairwise_sum(a[:n2]) + pairwise_sum(a[n2:]))
