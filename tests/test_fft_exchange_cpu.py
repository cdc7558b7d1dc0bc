"""The second owner exchange of the register-resident transform (csrc/fft_core.cuh) exists in two forms: through
shared memory (Owner::x2_write / x2_read) and through warp shuffles (exchange2_shuffle: log2(M2) rounds of
shfl.xor on G-point blocks + a register renaming).  This emulates both index maps on the CPU and checks that they
move every point to the same (thread, register) - for M2 = 2, 4, 8, i.e. N/2 or N = 512 ... 2048.  No GPU needed;
the GPU tests check the transforms themselves against torch.fft."""
import numpy as np
import pytest


def shared_memory_exchange(v, m2):
    """v[u][k] -> w[u][r] as written by x2_write(u, k) and read by x2_read(u, r), for one group of M2 threads (q = 0)."""
    g = 16 // m2
    words = {}
    for u in range(m2):  # j2 = u
        for k in range(16):
            words[u + m2 * k] = v[u][k]  # x2_write: j2 + M2 k
    w = np.empty_like(v)
    for u in range(m2):  # h = u
        for r in range(16):
            i, j2 = divmod(r, m2)
            w[u][r] = words[(u * g + i) * m2 + j2]  # x2_read: (h G + i) M2 + j2
    return w


def shuffle_exchange(v, m2):
    """exchange2_shuffle<M2>: every thread j2 runs the same rounds; shfl.xor pairs thread j2 with j2 ^ bit."""
    g = 16 // m2
    v = v.copy()
    bit = 1
    while bit < m2:
        send = np.empty((m2, 16), dtype=v.dtype)  # what each thread offers per (h0, i) slot of this round
        for j2 in range(m2):
            up = (j2 & bit) != 0
            for h0 in range(m2):
                if h0 & bit:
                    continue
                h1 = h0 | bit
                for i in range(g):
                    send[j2][h0 * g + i] = v[j2][h0 * g + i] if up else v[j2][h1 * g + i]
        for j2 in range(m2):
            up = (j2 & bit) != 0
            for h0 in range(m2):
                if h0 & bit:
                    continue
                h1 = h0 | bit
                for i in range(g):
                    recv = send[j2 ^ bit][h0 * g + i]
                    if up:
                        v[j2][h0 * g + i] = recv
                    else:
                        v[j2][h1 * g + i] = recv
        bit <<= 1
    w = np.empty_like(v)
    for j in range(m2):
        for i in range(g):
            w[:, i * m2 + j] = v[:, j * g + i]
    return w


@pytest.mark.parametrize("m2", [2, 4, 8])
def test_shuffle_exchange_equals_shared_memory_exchange(m2):
    v = np.arange(m2 * 16, dtype=np.int64).reshape(m2, 16) * 7 + 3  # distinct labels: (thread, register)
    a, b = shared_memory_exchange(v, m2), shuffle_exchange(v, m2)
    assert np.array_equal(a, b)
    assert sorted(a.ravel().tolist()) == sorted(v.ravel().tolist())  # a permutation of the points
