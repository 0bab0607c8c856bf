"""Generates tests/golden/xorwow_curand_host.npz from the REAL cuRAND library (host generator,
CURAND_RNG_PSEUDO_XORWOW, default ordering): output k*4096 + s is the k-th draw of the stream
curand_init(seed, subsequence = s, offset = 0) -- the call the reference makes
(lighttracing.cu:202-203).  Needs libcurand.so from the CUDA toolkit, no GPU.

    python tests/golden/make_golden.py
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def curand_host_uniforms(seed, n):
    cr = C.CDLL("libcurand.so")
    gen = C.c_void_p()
    assert cr.curandCreateGeneratorHost(C.byref(gen), 101) == 0  # CURAND_RNG_PSEUDO_XORWOW
    assert cr.curandSetPseudoRandomGeneratorSeed(gen, C.c_ulonglong(seed)) == 0
    out = np.empty(n, dtype=np.float32)
    assert cr.curandGenerateUniform(gen, out.ctypes.data_as(C.c_void_p), C.c_size_t(n)) == 0
    cr.curandDestroyGenerator(gen)
    return out


if __name__ == "__main__":
    seeds = np.array([0, 1, 7, 12345, 299999], dtype=np.uint32)
    data = np.stack([curand_host_uniforms(int(s), 4096 * 2) for s in seeds])
    np.savez_compressed(os.path.join(HERE, "xorwow_curand_host.npz"), seeds=seeds, uniforms=data)
    print("wrote", data.shape)
