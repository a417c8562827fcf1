"""Synthetic YUV 4:2:0 inputs, bit-identical to the generators inside oracle/ref_driver.c
(checked by tests/test_synth.py against `hl_ref_driver --dump-input`).

G1 "pan"    translating checkerboard + ramp + 2-bit noise (throughput runs)
G2 "stress" random base picture with half the pixels < 34 and 8x8 patches of 0 / 255, translated by (3n, -2n)
            (parity runs: both clip branches, the mod-256 wrap of the trial reconstruction, out-of-picture MVs)
"""
import numpy as np

_A = np.uint64(1664525)
_C = np.uint64(1013904223)
_M = np.uint64(0xFFFFFFFF)


def lcg_sequence(s0, n):
    """the n successive LCG states following s0 (s = s*1664525 + 1013904223 mod 2^32), vectorised by doubling"""
    if n <= 0:
        return np.zeros(0, np.uint32), np.uint32(s0)
    out = np.empty(n, np.uint64)
    out[0] = (np.uint64(s0) * _A + _C) & _M
    have, a_k, c_k = 1, _A, _C  # x[i+have] = a_k * x[i] + c_k
    while have < n:
        take = min(have, n - have)
        out[have:have + take] = (out[:take] * a_k + c_k) & _M
        c_k = (c_k * a_k + c_k) & _M
        a_k = (a_k * a_k) & _M
        have += take
    return out.astype(np.uint32), np.uint32(out[-1])


class G1:
    """frame n: Y = 128 + 60*(((x+2n)/8 + (y+n)/8)&1) + (((x+2n)*7 + (y+n)*13)&31) - 16 + (rnd()&3); UV[i] = 128 + ((i+n)&15)"""

    def __init__(self, width, height, seed=12345):
        self.w, self.h, self.state, self.n = width, height, np.uint32(seed), 0

    def next(self):
        w, h, n = self.w, self.h, self.n
        seq, self.state = lcg_sequence(self.state, w * h)
        noise = ((seq >> np.uint32(8)) & np.uint32(3)).astype(np.int32).reshape(h, w)
        x = np.arange(w, dtype=np.int32)[None, :] + 2 * n
        y = np.arange(h, dtype=np.int32)[:, None] + n
        v = 128 + 60 * (((x // 8) + (y // 8)) & 1) + ((x * 7 + y * 13) & 31) - 16 + noise
        yp = np.clip(v, 0, 255).astype(np.uint8)
        i = np.arange(w * h // 2, dtype=np.int32)
        uv = (128 + ((i + n) & 15)).astype(np.uint8)
        self.n += 1
        return np.concatenate([yp.reshape(-1), uv])


class G2:
    def __init__(self, width, height, seed=1):
        w, h = width, height
        self.w, self.h, self.n = w, h, 0
        s0 = np.uint32((np.uint64(seed) * np.uint64(2654435761) + np.uint64(97)) & _M)
        tot = w * h * 3 // 2
        seq, st = lcg_sequence(s0, 2 * tot)
        a = seq[0::2] >> np.uint32(8)
        b = seq[1::2] >> np.uint32(8)
        base = np.where((a & np.uint32(1)) == 1, b % np.uint32(34), b % np.uint32(256)).astype(np.uint8)
        nby, nbx = h // 8, w // 8
        seq2, _ = lcg_sequence(st, nby * nbx)
        sel = ((seq2 >> np.uint32(8)) & np.uint32(15)).reshape(nby, nbx)
        yb = base[:w * h].reshape(h, w).copy()
        for val, code in ((0, 0), (255, 1)):
            ys, xs = np.nonzero(sel == code)
            for by, bx in zip(ys, xs):
                yb[by * 8:by * 8 + 8, bx * 8:bx * 8 + 8] = val
        self.y = yb
        self.u = base[w * h:w * h * 5 // 4].reshape(h // 2, w // 2)
        self.v = base[w * h * 5 // 4:].reshape(h // 2, w // 2)

    def next(self):
        n, w, h = self.n, self.w, self.h
        # C: sx = ((x + 3n) % w + w) % w ; sy = ((y - 2n) % h + h) % h ; chroma: ((x + (3n)/2) % cw), ((y - n) % ch)
        yy = np.roll(np.roll(self.y, -(3 * n), axis=1), 2 * n, axis=0)
        cx = (3 * n) // 2
        uu = np.roll(np.roll(self.u, -cx, axis=1), n, axis=0)
        vv = np.roll(np.roll(self.v, -cx, axis=1), n, axis=0)
        self.n += 1
        return np.concatenate([yy.reshape(-1), uu.reshape(-1), vv.reshape(-1)])


class G3:
    """translating picture whose 8x8 blocks carry hashed texture of seven amplitudes over a slow ramp + 2-bit LCG noise (oracle/ref_driver.c:gen_g3): the
    block amplitudes straddle the homogeneity thresholds of the early-termination mode mask (rdo.c:889-935)"""
    AMP = np.array([0, 4, 10, 20, 36, 56, 90], np.uint64)

    def __init__(self, width, height, seed=1):
        self.w, self.h, self.seed, self.state, self.n = width, height, np.uint64(seed), np.uint32(12345), 0

    def next(self):
        w, h, n = self.w, self.h, self.n
        seq, self.state = lcg_sequence(self.state, w * h)
        noise = ((seq >> np.uint32(8)) & np.uint32(3)).astype(np.int64).reshape(h, w)
        X = (np.arange(w, dtype=np.uint64)[None, :] + np.uint64(2 * n))
        Y = (np.arange(h, dtype=np.uint64)[:, None] + np.uint64(n))
        v = (X * np.uint64(374761393) + Y * np.uint64(668265263) + self.seed * np.uint64(2246822519)) & _M
        v = ((v ^ (v >> np.uint64(13))) * np.uint64(1274126177)) & _M
        a = self.AMP[(((X >> np.uint64(3)) * np.uint64(73) + (Y >> np.uint64(3)) * np.uint64(151) + self.seed) & _M) % np.uint64(7)]
        s = 96 + (((X + Y) >> np.uint64(2)) & np.uint64(31)).astype(np.int64) + ((((v >> np.uint64(16)) & np.uint64(127)) * a) >> np.uint64(7)).astype(np.int64) + noise
        yp = np.minimum(s, 255).astype(np.uint8)
        i = np.arange(w * h // 2, dtype=np.int32)
        uv = (128 + ((i + n) & 15)).astype(np.uint8)
        self.n += 1
        return np.concatenate([yp.reshape(-1), uv])


def make(gen, width, height, seed=1):
    return G1(width, height) if gen == "g1" else (G3(width, height, seed) if gen == "g3" else G2(width, height, seed))
