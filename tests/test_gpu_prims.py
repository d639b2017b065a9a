"""The backend's own device-wide primitives on a real B200 (csrc/prims.cu replaces the reference's
kernels/radix_sort.cu + exclusive_scan.cu and the CUB calls of round 1): LSD radix sort of u32 key/value pairs and of u64
keys, unique of a sorted list -- bit-exact against std::stable_sort / std::unique on the host, from empty-ish and ragged
sizes (1, a partial tile, tile +- 1) up to the edge count of large-twist."""
import ctypes

import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,bits", [(1, 30), (2, 8), (31, 30), (4095, 30), (4096, 30), (4097, 22), (100003, 30), (1 << 20, 16),
                                    (3214258, 22), (9638218, 30)])
def test_radix_sort_and_unique_match_the_host(ppf, n, bits):
    lib = ctypes.CDLL(ppf.LIB_OURS)
    lib.ppf_b200_selftest_sort.restype = ctypes.c_int
    assert lib.ppf_b200_selftest_sort(ctypes.c_uint(n), ctypes.c_int(bits), ctypes.c_uint(n ^ 0x5EED)) == 0
