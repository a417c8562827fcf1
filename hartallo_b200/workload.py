"""Synthetic workloads for bench.py and the full-size parity tests (numpy only; no device code here).

"batch" workload = what the reference's encoder does to every macroblock of a P picture, expressed for the whole-frame batch
kernels: a list of ME candidates per MB (integer / half / quarter pattern points of me_ds.c:127-167 around the true motion,
for the 16x16, 16x8, 8x16 and 8x8 layouts), then a per-MB motion field (prediction) and the residual coding of the picture.
"""
import numpy as np

from . import lib as hl

# pattern points of hl_codec_264_me_ds_mb_find_best_cost (me_ds.c:127-167): 9 integer, 5 half, 9 quarter
_INT = [(0, 2), (-1, 1), (1, 1), (-2, 0), (0, 0), (2, 0), (-1, -1), (1, -1), (0, -2)]
_HALF = [(0, 1), (-1, 0), (0, -1), (1, 0), (0, 0)]
_QUARTER = [(-1, 1), (0, 1), (1, 1), (-1, 0), (0, 0), (1, 0), (-1, -1), (0, -1), (1, -1)]
PATTERN = np.array([(4 * x, 4 * y) for x, y in _INT] + [(2 * x, 2 * y) for x, y in _HALF] + list(_QUARTER), np.int16)  # quarter-pel offsets
# partitions (ox, oy, w, h) of the four macroblock layouts searched per MB
LAYOUTS = [(0, 0, 16, 16), (0, 0, 16, 8), (0, 8, 16, 8), (0, 0, 8, 16), (8, 0, 8, 16), (0, 0, 8, 8), (8, 0, 8, 8), (0, 8, 8, 8), (8, 8, 8, 8)]
CANDS_PER_MB = len(LAYOUTS) * len(PATTERN)  # 207
TRIALS_PER_MB = sum((w // 4) * (h // 4) for _, _, w, h in LAYOUTS) * len(PATTERN)  # 4x4 trial encodes per MB (1472)


def me_candidates(width, height, base_mv=(8, 4), seed=0):
    """ME_CAND array, CANDS_PER_MB per macroblock, MB-major; pattern centre = base_mv + a per-partition quarter-pel jitter"""
    mbw, mbh = width // 16, height // 16
    nmb = mbw * mbh
    rng = np.random.default_rng(seed)
    c = np.zeros((nmb, len(LAYOUTS), len(PATTERN)), hl.ME_CAND)
    mb = np.arange(nmb)
    c["mb_x"] = (mb % mbw)[:, None, None]
    c["mb_y"] = (mb // mbw)[:, None, None]
    for i, (ox, oy, w, h) in enumerate(LAYOUTS):
        c["part_x"][:, i], c["part_y"][:, i], c["part_w"][:, i], c["part_h"][:, i] = ox, oy, w, h
    jit = rng.integers(-3, 4, (nmb, len(LAYOUTS), 1, 2)).astype(np.int16)
    c["mv_x"] = base_mv[0] + jit[..., 0] + PATTERN[None, None, :, 0]
    c["mv_y"] = base_mv[1] + jit[..., 1] + PATTERN[None, None, :, 1]
    return c.reshape(-1)


def motion_field(width, height, base_mv=(8, 4), seed=1):
    """MB_MOTION array: random partition layout per MB, MVs = base_mv + quarter-pel jitter (all 16 fractional positions occur)"""
    nmb = (width // 16) * (height // 16)
    rng = np.random.default_rng(seed)
    m = np.zeros(nmb, hl.MB_MOTION)
    m["part_mode"] = rng.integers(0, 4, nmb)
    m["sub_mode"] = rng.integers(0, 4, (nmb, 4))
    m["mv"][..., 0] = base_mv[0] + rng.integers(-6, 7, (nmb, 4, 4))
    m["mv"][..., 1] = base_mv[1] + rng.integers(-6, 7, (nmb, 4, 4))
    return m


# integer-op accounting of SURVEY.md Appendix D: 560 ops per 4x4 trial + interpolation cost by fractional class
_FRAC_COST = np.zeros((4, 4), np.int64)  # [yf][xf]
for _yf in range(4):
    for _xf in range(4):
        if _xf == 0 and _yf == 0:
            _c = 0
        elif _yf == 0 or _xf == 0:
            _c = 176 if (_xf | _yf) == 2 else 208
        elif (_xf & 1) and (_yf & 1):
            _c = 352
        else:
            _c = 880
        _FRAC_COST[_yf, _xf] = _c


def me_int_ops(cands):
    """algorithmic 32-bit integer operations of costing `cands` (Appendix D)"""
    blocks = (cands["part_w"].astype(np.int64) // 4) * (cands["part_h"].astype(np.int64) // 4)
    frac = _FRAC_COST[cands["mv_y"].astype(np.int64) & 3, cands["mv_x"].astype(np.int64) & 3]
    return int((blocks * (560 + frac)).sum())


def cands_as_i32(cands):
    """n x 8 int32 view for oracle/hl_oracle.c:hlo_me_cost_batch"""
    out = np.zeros((len(cands), 8), np.int32)
    for k, f in enumerate(("mb_x", "mb_y", "part_x", "part_y", "part_w", "part_h", "mv_x", "mv_y")):
        out[:, k] = cands[f]
    return out


def motion_as_parts(motion):
    """nmb x 16 x 7 int32 {valid, ox, oy, w, h, mvx, mvy} for oracle/hl_oracle.c:hlo_predict_recon_mbs"""
    nmb = len(motion)
    out = np.zeros((nmb, 16, 7), np.int32)
    for mb in range(nmb):
        m = motion[mb]
        pm = int(m["part_mode"])
        if pm == 0:
            parts = [(0, 0, 0, 0, 16, 16)]
        elif pm == 1:
            parts = [(0, 0, 0, 0, 16, 8), (1, 0, 0, 8, 16, 8)]
        elif pm == 2:
            parts = [(0, 0, 0, 0, 8, 16), (1, 0, 8, 0, 8, 16)]
        else:
            parts = []
            for p in range(4):
                px, py, sm = (p & 1) * 8, (p >> 1) * 8, int(m["sub_mode"][p])
                if sm == 0:
                    parts.append((p, 0, px, py, 8, 8))
                elif sm == 1:
                    parts += [(p, 0, px, py, 8, 4), (p, 1, px, py + 4, 8, 4)]
                elif sm == 2:
                    parts += [(p, 0, px, py, 4, 8), (p, 1, px + 4, py, 4, 8)]
                else:
                    parts += [(p, s, px + (s & 1) * 4, py + (s >> 1) * 4, 4, 4) for s in range(4)]
        for k, (p, s, ox, oy, w, h) in enumerate(parts):
            out[mb, k] = (1, ox, oy, w, h, int(m["mv"][p, s, 0]), int(m["mv"][p, s, 1]))
    return out
