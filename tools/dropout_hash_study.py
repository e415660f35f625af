"""Statistical study for round 2: can the attention-dropout keep function get cheaper?  The tcgen05 attention kernels spend ~3.2 of their
~13 (forward) / ~25 (backward) instructions per probability on the hash (3 multiply-fold rounds per 4 keys, csrc/common.cuh
attn_drop_words) and the step is power-capped, so instructions are energy.  This script runs the battery of
tests/test_oracle_golden.py::test_attn_dropout_hash_statistics (keep rate, lag correlations along keys / rows / diagonals, spread of
per-row and per-key keep rates, cross-seed correlation, 2-D spectrum) on candidate generators, all keyed by the same per-row 64-bit key:

  r3      the current one: 3 rounds, 4 x 16-bit decisions per hash                       (13 instructions / 4 keys)
  r2      2 rounds (drop the third multiply), 4 x 16-bit decisions                        (~9 / 4 keys)
  r3x8    3 rounds, the 64 output bits cut into 8 x 8-bit decisions (p quantised to /256) (13 / 8 keys)
  r2x8    2 rounds, 8 x 8-bit decisions                                                   (~9 / 8 keys)
  r1      1 round, 4 x 16-bit decisions                                                   (~5 / 4 keys)  -- expected to fail

    python tools/dropout_hash_study.py
"""
import numpy as np

U = np.uint64
LO = U(0xFFFFFFFF)


def row_key(seed, rows):
    with np.errstate(over="ignore"):
        z = U(seed) + np.asarray(rows, dtype=U) * U(0x9E3779B97F4A7C15)
        z ^= z >> U(33); z *= U(0xFF51AFD7ED558CCD)
        z ^= z >> U(33); z *= U(0xC4CEB9FE1A85EC53)
        z ^= z >> U(33)
    return (z & LO)[:, None], (z >> U(32))[:, None]


def words(kind, k0, k1, g):
    with np.errstate(over="ignore"):
        a = (g * U(0x9E3779B1) + k0) & LO
        m1 = a * U(0x85EBCA6B)
        x = (m1 & LO) ^ (m1 >> U(32)) ^ k1
        if kind.startswith("r1"):
            return x, ((m1 >> U(32)) * U(0x9E3779B1) + (m1 & LO)) & LO
        m2 = x * U(0xC2B2AE35)
        if kind.startswith("r2"):
            return ((m2 >> U(32)) ^ (m1 & LO)) & LO, ((m2 & LO) ^ (m1 >> U(32)) ^ k0) & LO
        y = (m2 & LO) ^ (m2 >> U(32))
        m3 = y * U(0x27D4EB2F)
        return (m3 >> U(32)) ^ (m2 & LO), (m3 & LO) ^ (m2 >> U(32))


def keep_mask(kind, seed, rows, T, p):
    k0, k1 = row_key(seed, rows)
    if kind.endswith("x8"):
        g = np.arange((T + 7) // 8, dtype=U)[None, :]
        w0, w1 = words(kind, k0, k1, g)
        t8 = U(int(round(p * 256)))
        f = np.stack([(w >> U(s)) & U(0xFF) for w in (w0, w1) for s in (24, 16, 8, 0)], axis=-1)
        return (f >= t8).reshape(len(rows), -1)[:, :T], 1 - int(t8) / 256
    g = np.arange((T + 3) // 4, dtype=U)[None, :]
    w0, w1 = words(kind, k0, k1, g)
    t16 = int(round(p * 65536))
    t32 = U(t16 << 16)
    f = np.stack([w0, (w0 << U(16)) & LO, w1, (w1 << U(16)) & LO], axis=-1)
    return (f >= t32).reshape(len(rows), -1)[:, :T], 1 - t16 / 65536


def battery(kind):
    rows, T = np.arange(2048), 1156
    worst = dict(rate=0.0, lag=0.0, rowspread=0.0, colspread=0.0)
    for seed in (0, 7, 2 ** 63 + 12345):
        m, q = keep_mask(kind, seed, rows, T, 0.1)
        m = m.astype(np.float64); n = m.size; var = q * (1 - q)
        worst["rate"] = max(worst["rate"], abs(m.mean() - q) / np.sqrt(var / n))                       # in sigmas (limit 4)
        c = m - m.mean(); v = (c * c).mean()
        for (dr, dk) in ((0, 1), (0, 2), (0, 3), (0, 4), (0, 5), (0, 7), (0, 8), (1, 0), (1, 1), (2, 0), (16, 0), (1, 4), (1, 8)):
            r = (c[dr:, dk:] * c[:c.shape[0] - dr, :c.shape[1] - dk]).mean() / v
            worst["lag"] = max(worst["lag"], abs(r) * np.sqrt(n))                                       # in sigmas (limit 5)
        worst["rowspread"] = max(worst["rowspread"], abs(m.mean(1).std() / np.sqrt(var / T) - 1))        # limit 0.10
        worst["colspread"] = max(worst["colspread"], abs(m.mean(0).std() / np.sqrt(var / len(rows)) - 1))  # limit 0.15
    a, q = keep_mask(kind, 5, rows, T, 0.1); b, _ = keep_mask(kind, 6, rows, T, 0.1)
    a = a.astype(np.float64); b = b.astype(np.float64)
    cross = abs(((a - a.mean()) * (b - b.mean())).mean() / (q * (1 - q))) * np.sqrt(a.size)               # limit 5
    h, _ = keep_mask(kind, 9, np.arange(1024), 1024, 0.5)
    h = h.astype(np.float64) - 0.5
    spec = np.abs(np.fft.fft2(h)) ** 2 / (1024 * 1024 * 0.25)
    spec[0, 0] = 0
    ok = (worst["rate"] < 4 and worst["lag"] < 5 and worst["rowspread"] < 0.10 and worst["colspread"] < 0.15 and cross < 5 and spec.max() < 25
          and abs(spec.mean() - 1) < 0.01)
    return ok, worst, cross, spec.max(), spec.mean()


if __name__ == "__main__":
    print("%-6s %-5s %8s %8s %10s %10s %8s %9s %9s" % ("kind", "pass", "rate/s", "lag/s", "rowspread", "colspread", "cross/s", "spec max", "spec mean"))
    for kind in ("r3", "r2", "r3x8", "r2x8", "r1"):
        ok, w, cross, smax, smean = battery(kind)
        print("%-6s %-5s %8.2f %8.2f %10.3f %10.3f %8.2f %9.1f %9.4f" % (kind, "yes" if ok else "NO", w["rate"], w["lag"], w["rowspread"], w["colspread"],
                                                                       cross, smax, smean))
