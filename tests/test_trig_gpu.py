"""GPU: the device's cosf/sinf (csrc/trig_glibc.cuh, a port of glibc's sincosf algorithm with the FMA build's
contraction pattern) against the host libm for EVERY float angle in [0, 360] degrees (1.136e9 values) —
the a,b of computeOrbDescriptor (src/ORBextractor.cc:108-109) must be bit-identical or descriptors drift."""
import ctypes as C
from concurrent.futures import ThreadPoolExecutor
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_cos_sin_bit_exact_all_angles(gpu, oracle):
    L = gpu.lib()
    last = int(np.float32(360.0).view(np.uint32))
    chunk = 1 << 24
    starts = list(range(0, last + 1, chunk))

    def host(first, n):
        deg = (np.arange(first, first + n, dtype=np.uint32)).view(np.float32)
        return oracle.cos_sin_deg(deg)

    bad = 0
    with ThreadPoolExecutor(8) as pool:
        futs = {}
        nxt = 0
        for k, st in enumerate(starts):
            while nxt < len(starts) and nxt < k + 8:             # at most 8 host chunks in flight (64 MB x 2 each)
                s0 = starts[nxt]
                futs[s0] = pool.submit(host, s0, min(chunk, last + 1 - s0))
                nxt += 1
            n = min(chunk, last + 1 - st)
            c = np.empty(n, np.float32)
            s = np.empty(n, np.float32)
            gpu._check(L.ygz_debug_cos_sin(0, st, n, c.ctypes.data, s.ctypes.data), "ygz_debug_cos_sin")
            hc, hs = futs.pop(st).result()
            bad += int((c.view(np.uint32) != hc.view(np.uint32)).sum()) + int((s.view(np.uint32) != hs.view(np.uint32)).sum())
    assert bad == 0, "%d of %d (cos, sin) values differ from the host libm" % (bad, 2 * (last + 1))
