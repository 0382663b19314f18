"""Deterministic synthetic NV12 / P010 frames — the inputs of tests and bench.py.

A scene is a function of the frame index t: a multi-octave value-noise background that translates by
(+6, -3) luma pixels per frame, three textured rectangles (W/8 x H/8) moving by (+17, 0), (0, -11) and
(-9, +9) px per frame, plus +-2 uniform noise.  Everything derives from SplitMix64 hashes of
(seed, field, lattice coordinates), so frames are reproducible anywhere and any t can be generated
independently.  P010: 10-bit value = 4 * (8-bit field) + 2 noise bits, stored in the 10 MSBs.
"""
import numpy as np

SEED_BASE = 0x4852423230300000  # "HRB200\0\0"
GLOBAL_MOTION = (6, -3)
RECT_MOTIONS = ((17, 0), (0, -11), (-9, 9))

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix(z):
    """SplitMix64 finaliser on a uint64 array."""
    with np.errstate(over="ignore"):
        z = (z + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def _hash(seed, field, i, j):
    with np.errstate(over="ignore"):
        k = (np.uint64(seed) + np.uint64(field) * np.uint64(0xD6E8FEB86659FD93)) & _M64
        k = _splitmix(k ^ (i.astype(np.int64).astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)))
        k = _splitmix(k ^ (j.astype(np.int64).astype(np.uint64) * np.uint64(0xC2B2AE3D27D4EB4F)))
    return k


def _unit(h):
    """uint64 hash -> float64 in [-1, 1)."""
    return (h >> np.uint64(11)).astype(np.float64) * (2.0 / (1 << 53)) - 1.0


def _value_noise(seed, field, xs, ys, period):
    """Bilinear value noise on an infinite lattice; xs (1-D, columns) and ys (1-D, rows) in luma pixels."""
    fx = xs / float(period)
    fy = ys / float(period)
    ix = np.floor(fx).astype(np.int64)
    iy = np.floor(fy).astype(np.int64)
    tx = (fx - ix)[None, :]
    ty = (fy - iy)[:, None]
    ux, inv_x = np.unique(np.concatenate([ix, ix + 1]), return_inverse=True)
    uy, inv_y = np.unique(np.concatenate([iy, iy + 1]), return_inverse=True)
    lat = _unit(_hash(seed, field, ux[None, :].repeat(len(uy), 0), uy[:, None].repeat(len(ux), 1)))
    n = len(ix)
    m = len(iy)
    x0, x1 = inv_x[:n], inv_x[n:]
    y0, y1 = inv_y[:m], inv_y[m:]
    top = lat[np.ix_(y0, x0)] * (1 - tx) + lat[np.ix_(y0, x1)] * tx
    bot = lat[np.ix_(y1, x0)] * (1 - tx) + lat[np.ix_(y1, x1)] * tx
    return top * (1 - ty) + bot * ty


def _field(seed, base_field, xs, ys, octaves):
    acc = np.zeros((len(ys), len(xs)))
    for k, (period, amp) in enumerate(octaves):
        acc += amp * _value_noise(seed, base_field + k, xs, ys, period)
    return acc


_LUMA_OCT = ((64, 96.0), (32, 48.0), (16, 24.0), (8, 12.0))
_CHROMA_OCT = ((64, 48.0), (32, 24.0))


def _scene_plane(seed, W, H, t, xs, ys, kind):
    """kind 0: luma, 1: U, 2: V.  xs/ys are the luma-pixel coordinates of the samples."""
    octs = _LUMA_OCT if kind == 0 else _CHROMA_OCT
    gx, gy = GLOBAL_MOTION
    img = 128.0 + _field(seed, 16 * kind, xs - gx * t, ys - gy * t, octs)
    rw, rh = max(W // 8, 2), max(H // 8, 2)
    starts = ((W // 8, H // 6), (W // 2, (2 * H) // 3), ((3 * W) // 4, H // 3))
    for r, ((vx, vy), (sx, sy)) in enumerate(zip(RECT_MOTIONS, starts)):
        x0, y0 = sx + vx * t, sy + vy * t
        cx = np.nonzero((xs >= x0) & (xs < x0 + rw))[0]
        cy = np.nonzero((ys >= y0) & (ys < y0 + rh))[0]
        if len(cx) == 0 or len(cy) == 0:
            continue
        tex = 128.0 + _field(seed, 100 + 16 * kind + 4 * r, xs[cx] - x0, ys[cy] - y0, octs) * 0.9 + (18.0 if kind == 0 else 0.0) * (r - 1)
        img[np.ix_(cy, cx)] = tex
    return img


def _pixel_noise(seed, field, t, w, h, lo, hi):
    xs = np.arange(w, dtype=np.int64)[None, :].repeat(h, 0)
    ys = np.arange(h, dtype=np.int64)[:, None].repeat(w, 1)
    hsh = _hash(seed + 7919 * (t + 1), field, xs, ys)
    return (hsh % np.uint64(hi - lo + 1)).astype(np.int64) + lo


def frame_bytes(width, height, hdr, stride=None):
    stride = stride or width
    return (height * stride + (height // 2) * stride) * (2 if hdr else 1)


def make_frame(width, height, t, seed=SEED_BASE, hdr=False, stride=None, noise=True):
    """Frame t of the synthetic scene as a flat array (uint8 NV12 or uint16 P010) of 1.5*H*stride elements."""
    stride = stride or width
    xs = np.arange(width, dtype=np.float64)
    ys = np.arange(height, dtype=np.float64)
    luma = _scene_plane(seed, width, height, t, xs, ys, 0)
    cxs = np.arange(width // 2, dtype=np.float64) * 2.0
    cys = np.arange(height // 2, dtype=np.float64) * 2.0
    u = _scene_plane(seed, width, height, t, cxs, cys, 1)
    v = _scene_plane(seed, width, height, t, cxs, cys, 2)
    if noise:
        luma = luma + _pixel_noise(seed, 900, t, width, height, -2, 2)
        u = u + _pixel_noise(seed, 901, t, width // 2, height // 2, -1, 1)
        v = v + _pixel_noise(seed, 902, t, width // 2, height // 2, -1, 1)
    y8 = np.clip(np.rint(luma), 16, 235).astype(np.int64)
    u8 = np.clip(np.rint(u), 16, 240).astype(np.int64)
    v8 = np.clip(np.rint(v), 16, 240).astype(np.int64)
    dtype = np.uint16 if hdr else np.uint8
    out = np.zeros((height + height // 2, stride), dtype)
    if hdr:
        y10 = y8 * 4 + _pixel_noise(seed, 910, t, width, height, 0, 3)
        u10 = u8 * 4 + _pixel_noise(seed, 911, t, width // 2, height // 2, 0, 3)
        v10 = v8 * 4 + _pixel_noise(seed, 912, t, width // 2, height // 2, 0, 3)
        out[:height, :width] = (y10 << 6).astype(dtype)
        out[height:, 0:width:2] = (u10 << 6).astype(dtype)
        out[height:, 1:width:2] = (v10 << 6).astype(dtype)
    else:
        out[:height, :width] = y8.astype(dtype)
        out[height:, 0:width:2] = u8.astype(dtype)
        out[height:, 1:width:2] = v8.astype(dtype)
    if stride > width:  # padding bytes are deterministic garbage: results must not depend on them
        pad = _pixel_noise(seed, 990, t, stride - width, height + height // 2, 0, 255 if not hdr else 65535)
        out[:, width:] = pad.astype(dtype)
    return out.reshape(-1)


def make_random_frame(width, height, seed, hdr=False, stride=None):
    """Uniform random samples: forces uint32 wrap-around of large window sums and arg-min near-ties."""
    stride = stride or width
    rng = np.random.Generator(np.random.PCG64(seed))
    if hdr:
        return (rng.integers(0, 1024, (height + height // 2) * stride, dtype=np.uint16) << 6).astype(np.uint16)
    return rng.integers(0, 256, (height + height // 2) * stride, dtype=np.uint8)


def make_ramp_frame(width, height, shift=0, hdr=False, stride=None):
    """Horizontal luma ramp / vertical chroma ramp (exercises the mirrored borders)."""
    stride = stride or width
    dtype = np.uint16 if hdr else np.uint8
    out = np.zeros((height + height // 2, stride), dtype)
    x = (np.arange(width) + shift) % 256
    y = np.arange(height // 2) % 256
    out[:height, :width] = x[None, :]
    out[height:, :width] = y[:, None]
    if hdr:
        out = (out.astype(np.uint32) << 8).astype(np.uint16)
    return out.reshape(-1)
