"""CPU model of the shared-memory addressing of K-COL's staged two-column tiles at NX = 8192 (csrc/xfb_colt.cuh,
ColTCfg::SWZ): the tiles are moved by TMA with CU_TENSOR_MAP_SWIZZLE_32B (byte address bit 4 ^= bit 7) and every thread
applies the same exchange to its column chunk (`swz = (t >> 3) & 1`).  Checked here, without a GPU:
  * the thread-side formula is exactly the TMA pattern, for every row and both columns;
  * it is a bijection of the tile (nothing lost, nothing overwritten);
  * a half warp's 8-byte accesses (one wavefront of 128 bytes) fall on 16 distinct bank pairs with the swizzle and on
    8 without (the 2-way conflict the swizzle removes)."""
import numpy as np

NX, TW, G = 8192, 2, 512          # rows, columns per tile, butterfly threads (row i = t + k * G)


def logical_byte(i, col):
    """dense TMA box layout: [row pair][column][row parity], 8-byte elements"""
    return (((i >> 1) * TW + col) * 2 + (i & 1)) * 8


def tma_swizzle_32b(byte):
    return byte ^ (((byte >> 7) & 1) << 4)


def thread_byte(t, k, col, swizzle=True):
    """what colt_kernel computes: S + s_base + 2 * (col ^ swz) + k * G * TW"""
    swz = ((t >> 3) & 1) if swizzle else 0
    s_base = ((t >> 1) * TW) * 2 + (t & 1)
    return (s_base + 2 * (col ^ swz) + k * G * TW) * 8


def test_thread_formula_is_the_tma_pattern():
    t = np.arange(G)
    for k in range(16):
        for col in range(TW):
            i = t + k * G
            assert np.array_equal(thread_byte(t, k, col), tma_swizzle_32b(logical_byte(i, col)))


def test_swizzle_is_a_bijection_of_the_tile():
    t, k, col = np.meshgrid(np.arange(G), np.arange(16), np.arange(TW), indexing="ij")
    b = thread_byte(t, k, col).ravel()
    assert b.min() == 0 and b.max() == NX * TW * 8 - 8 and np.unique(b).size == NX * TW


def _bank_pairs(bytes_):
    return np.unique((bytes_ % 128) // 8).size


def test_half_warp_accesses_are_conflict_free_only_with_the_swizzle():
    for warp in range(G // 32):
        for half in range(2):
            t = np.arange(16) + 32 * warp + 16 * half
            for k in (0, 5, 15):
                for col in range(TW):
                    assert _bank_pairs(thread_byte(t, k, col, True)) == 16
                    assert _bank_pairs(thread_byte(t, k, col, False)) == 8


def test_ring_box_reads_use_the_same_exchange():
    """ColtRing::step reads column 0 at src[c0_off] and column 1 at src[2 - c0_off], c0_off = 2 * swz, relative to
    s_off = s_base (elements of 8 bytes) inside a box of 256 row pairs"""
    t = np.arange(G)
    swz = (t >> 3) & 1
    s_base = ((t >> 1) * TW) * 2 + (t & 1)
    col0 = (s_base + 2 * swz) * 8
    col1 = (s_base + 2 - 2 * swz) * 8
    assert np.array_equal(col0, tma_swizzle_32b(logical_byte(t, 0)))
    assert np.array_equal(col1, tma_swizzle_32b(logical_byte(t, 1)))
