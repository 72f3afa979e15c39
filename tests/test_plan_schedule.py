"""CPU-only: the host-side MMA schedule (csrc/plan.cpp) replayed in NumPy must reproduce the oracle.

This exercises exactly what the tensor-core kernel will do - 16-padded K slots, K=32 windows, pairs of
adjacent blocks, chunk boundaries, block-row groups - without needing a GPU."""
import ctypes as C

import numpy as np
import pytest

from oracle import bsr_oracle as O
from resnet_accel_b200 import _lib

CHUNK = 9


def _export(rp, ci, nbr, nbc, group_rows=0):
    L = _lib.lib()
    h, ws = C.c_void_p(), C.c_size_t()
    rp = np.ascontiguousarray(rp, np.int32)
    ci = np.ascontiguousarray(ci, np.int32)
    rc = L.accel_plan_create(rp.ctypes.data, ci.ctypes.data if ci.size else None, nbr, nbc, 14, group_rows,
                             C.byref(h), C.byref(ws))
    assert rc == 0, L.accel_last_error_string()
    n = L.accel_plan_export_ops(h, None, 0)
    rec = np.zeros((max(n, 1), 8), np.int32)
    L.accel_plan_export_ops(h, rec.ctypes.data, n)
    nblk = L.accel_plan_num_blocks(h)
    nm = L.accel_plan_export_mma(h, None, 0)
    mma = np.zeros((max(nm, 1), 8), np.int32)
    L.accel_plan_export_mma(h, mma.ctypes.data, nm)
    assert nm == L.accel_plan_num_mma(h) and n == L.accel_plan_num_tiles(h)
    L.accel_plan_destroy(h)
    _check_mma_cover(rec[:n], mma[:nm])
    return rec[:n], nblk, ws.value


def _check_mma_cover(rec, mma):
    """Every weight tile is covered by exactly one tcgen05.mma record, and a wide MMA (N = 16*len) only merges
    tiles of adjacent block-rows that share group, batch, K chunk and window."""
    covered = np.zeros(len(rec), np.int32)
    for grp, batch, dcol, N, acol, tile0, chunk, br0 in mma:
        assert N % 16 == 0 and 16 <= N <= 176 and dcol % 16 == 0 and dcol + N <= 176
        assert acol % 4 == 0 and 0 <= acol // 4 <= CHUNK - 2
        for i in range(N // 16):
            t = rec[tile0 + i]
            assert (t[0], t[1], t[2], t[3], t[4], t[7]) == (grp, br0, dcol // 16 + i, chunk, acol // 4, batch)
            covered[tile0 + i] += 1
    assert np.all(covered == 1)
    assert len(mma) <= len(rec)


def _replay(A, rec, blocks, nbr, nbc):
    M, K = A.shape
    n_chunks = -(-nbc // CHUNK)
    X16 = np.zeros((M, n_chunks * CHUNK + 1, 16), np.int64)      # 16-byte K slots (14 data + 2 zero)
    Ap = np.zeros((M, (n_chunks * CHUNK + 1) * 14), np.int64)
    Ap[:, :K] = A
    X16[:, :, :14] = Ap.reshape(M, -1, 14)
    Y = np.zeros((M, nbr * 14), np.int64)
    for grp, br0, g, chunk, win, lo, hi, _ in rec:
        assert 0 <= win <= CHUNK - 2
        t = chunk * CHUNK + win
        window = X16[:, t:t + 2, :].reshape(M, 32)
        tile = np.zeros((16, 32), np.int64)
        if lo >= 0:
            tile[:14, 0:14] = blocks[lo]
        if hi >= 0:
            tile[:14, 16:30] = blocks[hi]
        Y[:, (br0 + g) * 14:(br0 + g + 1) * 14] += (window @ tile.T)[:, :14]
    return Y.astype(np.int32)


@pytest.mark.parametrize("N,K,density,group_rows", [
    (14, 14, 1.0, 0), (140, 9226, 1.0, 0), (500, 1000, 0.3, 0), (1000, 512, 0.7, 8), (449, 3000, 0.0, 0),
    (448, 2304, 1.0, 5), (15, 225, 0.6, 0), (600, 237, 0.9, 32), (64, 147, 0.3, 0), (4102, 500, 0.5, 0),
])
def test_schedule_replay_matches_oracle(N, K, density, group_rows):
    rng = np.random.default_rng(N + K)
    W = rng.integers(-128, 128, (N, K), dtype=np.int8)
    nbr, nbc = -(-N // 14), -(-K // 14)
    keep = rng.random((nbr, nbc)) < density
    W = W * np.repeat(np.repeat(keep, 14, 0), 14, 1)[:N, :K].astype(np.int8)
    bsr = O.build_bsr_14x14_int8_direct(W)
    rec, nblk, ws = _export(bsr["indptr"], bsr["indices"], nbr, nbc, group_rows)
    assert nblk == bsr["num_blocks"]
    # every stored block is used exactly once
    used = np.concatenate([rec[:, 5], rec[:, 6]])
    used = np.sort(used[used >= 0])
    assert np.array_equal(used, np.arange(nblk))
    # ops of one batch share a chunk; batches are visited in order
    assert np.all(np.diff(rec[:, 7]) >= 0)
    A = rng.integers(-128, 128, (3, K), dtype=np.int8)
    ref = O.bsr_gemm_i32(A, bsr["indptr"], bsr["indices"], bsr["data"])
    assert np.array_equal(_replay(A, rec, bsr["data"], nbr, nbc), ref)
    # pairing efficiency: never more MMAs than blocks, at least half
    assert (nblk + 1) // 2 <= len(rec) <= max(nblk, 0)


def test_header_symbols_exported():
    """Every function include/accel_b200.h declares is exported by the .so and bound in _lib.SYMBOLS."""
    import os
    import re
    hdr = open(os.path.join(os.path.dirname(_lib.LIB_PATH), "..", "include", "accel_b200.h")).read()
    declared = set(re.findall(r"ACCEL_API [\w\s\*]+?\b(accel_\w+)\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name)


def test_structure_validation_messages():
    L = _lib.lib()
    h, ws = C.c_void_p(), C.c_size_t()
    rp = np.array([0, 2], np.int32)
    for ci, frag in (([1, 0], b"not sorted"), ([0, 9], b"exceeds num_block_cols")):
        ci = np.array(ci, np.int32)
        assert L.accel_plan_create(rp.ctypes.data, ci.ctypes.data, 1, 2, 14, 0, C.byref(h), C.byref(ws)) == _lib.INVALID_CONFIG
        assert frag in L.accel_last_error_string()
    ci = np.array([0, 1], np.int32)
    assert L.accel_plan_create(rp.ctypes.data, ci.ctypes.data, 1, 2, 8, 0, C.byref(h), C.byref(ws)) == _lib.INVALID_CONFIG
    assert b"Block size must be 14" in L.accel_last_error_string()
