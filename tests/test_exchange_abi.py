"""Host-side checks of the multi-GPU exchange entry points (no GPU needed: sizes and argument validation
happen before any CUDA call).  The kernels themselves are covered by the -m gpu tests
(test_gpu_parity.py::test_push_exchange_kernels_on_one_device, test_multi_rank_gpu.py)."""
import ctypes as C

import pytest

from pollen_b200 import binding as pb


def _per(n_ranks, n_segs):
    n_words = (n_segs + 31) // 32
    return max((((n_words + n_ranks - 1) // n_ranks) + 31) // 32 * 32, 32)


@pytest.mark.parametrize("n_ranks", [1, 2, 3, 4, 7, 8, 16])
@pytest.mark.parametrize("n_segs", [1, 31, 32, 33, 4096, 1_000_000, 5_000_000, 5_000_001])
def test_receive_buffer_holds_one_slot_per_rank_of_whole_cache_lines(n_ranks, n_segs):
    nbytes = pb.exchange_recv_bytes(n_ranks, n_segs)
    per = _per(n_ranks, n_segs)
    assert nbytes == n_ranks * per * 160                 # a slot = per words x 32 segments x (u32 depth + u8 uniq)
    assert per % 32 == 0 and per * n_ranks * 32 >= n_segs  # the slices cover the segment axis
    assert nbytes % 128 == 0


def test_exchange_entry_points_reject_bad_arguments_before_touching_the_device():
    lib = pb.lib()
    P2 = C.c_void_p * 2
    bufs = P2(0x1000, 0x2000)
    inv = -1                                              # FGFA_ERR_INVALID_ARG
    assert pb.exchange_recv_bytes(0, 1000) == 0
    # rank out of range, too many ranks, null receive buffers, misaligned receive buffer, > 255 rows
    assert lib.fgfa_exchange_push(2, 2, None, 0, 0x1000, None, bufs, 64, None) == inv
    assert lib.fgfa_exchange_push(17, 0, None, 0, 0x1000, None, bufs, 64, None) == inv
    assert lib.fgfa_exchange_push(2, 0, None, 0, 0x1000, None, None, 64, None) == inv
    assert lib.fgfa_exchange_push(2, 0, None, 0, 0x1000, None, P2(0x1008, 0x2000), 64, None) == inv
    assert lib.fgfa_exchange_push(2, 0, 0x3000, 256, 0x1000, None, bufs, 64, None) == inv
    assert lib.fgfa_exchange_push(2, 0, None, 3, 0x1000, None, bufs, 64, None) == inv      # rows without a bitmap
    assert lib.fgfa_exchange_reduce(2, -1, 0x1000, bufs, bufs, 64, None, 0, 0, None) == inv
    assert lib.fgfa_exchange_reduce(2, 0, None, bufs, bufs, 64, None, 0, 0, None) == inv
    assert lib.fgfa_exchange_reduce(2, 0, 0x1000, bufs, bufs, 64, 0x4000, 8, 0, None) == inv   # multicast offsets: 16-byte aligned
    assert b"exchange" in lib.fgfa_last_error() or b"multicast" in lib.fgfa_last_error()
    # an empty graph is a no-op
    assert lib.fgfa_exchange_push(2, 0, None, 0, None, None, bufs, 0, None) == 0
    assert lib.fgfa_exchange_reduce(2, 0, 0x1000, bufs, bufs, 0, None, 0, 0, None) == 0
