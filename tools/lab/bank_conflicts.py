#!/usr/bin/env python
"""Enumerates the shared-memory bank conflicts of the x pass's exchange patterns (csrc/fft_core.cuh: LineAddr / line_skew)
for 8-byte exchange words: a warp access is served in two phases of 16 lanes; a phase is conflict-free when its 16 words
fall into 16 different 8-byte bank pairs (word index mod 16).  Prints, per grid size and access pattern, the worst and the
mean number of wavefronts per phase with and without the skew.  Run on the CPU; no GPU needed."""


def analyze(logn, skew):
    n = 1 << logn
    m1 = n // 16
    m2 = max(m1 // 16, 1)
    g = 16 // m2

    def x1_write(u, q):
        return q * m1 + u

    def x1_read(u, k):
        q, j2 = divmod(u, m2)
        return q * m1 + j2 + m2 * k

    def x2_read(u, r):
        q, h = divmod(u, m2)
        i, j2 = divmod(r, m2)
        return q * m1 + (h * g + i) * m2 + j2

    def out_freq(u, r):
        q, h = divmod(u, m2)
        i, k3 = divmod(r, m2)
        return q + 16 * (h * g + i) + 256 * k3

    def split_read(u, r):  # Z[k] for r < 8, Z[N - k] for r >= 8, k = u + M1 (r mod 8)
        k = u + m1 * (r % 8)
        return k if r < 8 else (n - k) % n

    patterns = {"exchange 1 write": x1_write, "exchange 1 read / exchange 2 write": x1_read, "exchange 2 read": x2_read,
                "split write": out_freq, "split read": split_read}
    assert len({skew(e) for e in range(n)}) == n, "the skew must be injective"
    res = {}
    for name, f in patterns.items():
        worst, total, count = 0, 0, 0
        for reg in range(16):
            for w0 in range(0, m1, 16):
                banks = {}
                for u in range(w0, min(w0 + 16, m1)):
                    a = skew(f(u, reg))
                    banks.setdefault(a % 16, set()).add(a)
                deg = max(len(v) for v in banks.values())
                worst, total, count = max(worst, deg), total + deg, count + 1
        res[name] = (worst, round(total / count, 2))
    return res


def line_skew(e):
    return e + 4 * (e >> 6) + ((e >> 4) & 3)


if __name__ == "__main__":
    for logn in (8, 9, 10, 11):
        print(f"N = {1 << logn}")
        plain = analyze(logn, lambda e: e)
        skewed = analyze(logn, line_skew)
        for name in plain:
            print(f"  {name:38s} no skew: worst {plain[name][0]:2d} mean {plain[name][1]:5.2f}   line_skew: worst "
                  f"{skewed[name][0]} mean {skewed[name][1]}")
