"""GPU check of the hand-written radix sort (csrc/radix_sort.cu) against numpy's stable sort.

The sort is the grouping step of apply_gradients: (slot, batch index) pairs, stable, on the low `end_bit` key bits.
It is reached through a test hook that libmeepo.so exports outside include/meepo.h.
"""
import ctypes as C

import numpy as np
import pytest

from meepoembedding_b200 import Table

from util import table_kwargs

pytestmark = pytest.mark.gpu


def _sort(t, keys, vals, end_bit):
    import torch

    fn = t.lib.dll.meepo_internal_sort_pairs
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int32, C.c_void_p]
    dk = torch.from_numpy(keys.view(np.int32)).cuda()
    dv = torch.from_numpy(vals.view(np.int32)).cuda()
    ok, ov = torch.empty_like(dk), torch.empty_like(dv)
    t.lib.check(fn(t._h, dk.data_ptr(), ok.data_ptr(), dv.data_ptr(), ov.data_ptr(), keys.size, end_bit, None))
    torch.cuda.synchronize()
    return ok.cpu().numpy().view(np.uint32), ov.cpu().numpy().view(np.uint32)


@pytest.mark.parametrize("n", [1, 31, 32, 33, 4095, 4096, 4097, 100_003, 3_000_001])
@pytest.mark.parametrize("end_bit", [1, 7, 8, 9, 16, 22, 28, 32])
def test_radix_sort_matches_stable_numpy(cuda_lib, n, end_bit):
    t = Table(lib=cuda_lib, **table_kwargs(dim=4, capacity=64))
    rng = np.random.default_rng([n, end_bit])
    hi = (1 << end_bit) - 1
    for kind in ("random", "few_values", "constant"):
        if kind == "random":
            keys = rng.integers(0, hi + 1, size=n, dtype=np.uint64).astype(np.uint32)
        elif kind == "few_values":
            keys = rng.choice(rng.integers(0, hi + 1, size=5, dtype=np.uint64), size=n).astype(np.uint32)
        else:
            keys = np.full(n, hi, dtype=np.uint32)
        # bits above end_bit must be ignored by the sort (and carried along)
        junk = (rng.integers(0, 1 << 32, size=n, dtype=np.uint64) << np.uint64(end_bit)).astype(np.uint32) if end_bit < 32 else 0
        full = keys | junk
        vals = np.arange(n, dtype=np.uint32)
        gk, gv = _sort(t, full, vals, end_bit)
        order = np.argsort(keys, kind="stable")
        np.testing.assert_array_equal(gv, vals[order], err_msg=f"{kind}")
        np.testing.assert_array_equal(gk, full[order], err_msg=f"{kind}")
    t.stats()  # raises if the look-back ever gave up
    t.close()
