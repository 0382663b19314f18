"""CPU tests of the oracle itself (the checker must be right before it checks anything)."""
import numpy as np
import pytest

from hopperrender_b200 import synth
from oracle import OracleCalc, kernels


def sq(d):
    return d * d * (1 if d > 0 else -1)


def mirror(n, dim):
    if n >= dim:
        n = dim - (n - dim + 1)
    elif n < 0:
        n = -n - 1
    return min(max(n, 0), dim - 1)


def closed_form_sums(f1, f2, offs, H, W, S, ws, R, rs, iteration, step, ds, ns, hdr):
    """SURVEY.md A.1 per-window closed form, in plain python/numpy (independent of the C++ oracle):
    sums[z][w] = (SAD << ds) + n_w * (|o + sq(d)| + (NB << ns))  mod 2^32."""
    _, lh, lw = offs.shape
    sh = 8 if hdr else 0
    Y1 = (f1[:H * S].reshape(H, S) >> sh).astype(np.int64)
    Y2 = (f2[:H * S].reshape(H, S) >> sh).astype(np.int64)
    C1 = (f1[H * S:H * S + (H // 2) * S].reshape(H // 2, S) >> sh).astype(np.int64)
    C2 = (f2[H * S:H * S + (H // 2) * S].reshape(H // 2, S) >> sh).astype(np.int64)
    nWy, nWx = -(-lh // ws), -(-lw // ws)
    out = np.zeros((R, nWy, nWx), np.uint64)
    for wy in range(nWy):
        for wx in range(nWx):
            y0, x0 = wy * ws, wx * ws
            ys = np.arange(y0, min(y0 + ws, lh))
            xs = np.arange(x0, min(x0 + ws, lw))
            ox, oy = int(offs[0, y0, x0]), int(offs[1, y0, x0])
            assert (offs[0, ys[0]:ys[-1] + 1, xs[0]:xs[-1] + 1] == ox).all()
            nw = len(ys) * len(xs)
            for z in range(R):
                d = sq(z - R // 2)
                cx, cy = (ox + d, oy) if step == 0 else (ox, oy + d)
                sy, sx = ys << rs, xs << rs
                ny = np.array([mirror(int(v) + cy, H) for v in sy])
                nx = np.array([mirror(int(v) + cx, W) for v in sx])
                sad = np.abs(Y1[np.ix_(ny, nx)] - Y2[np.ix_(sy, sx)]).sum()
                sad += np.abs(C1[np.ix_(ny >> 1, nx & ~1)] - C2[np.ix_(sy >> 1, sx & ~1)]).sum()
                sad += np.abs(C1[np.ix_(ny >> 1, (nx & ~1) + 1)] - C2[np.ix_(sy >> 1, (sx & ~1) + 1)]).sum()
                cand = cx if step == 0 else cy
                bias = abs(cand)
                if iteration >= 4:
                    pl = offs[0] if step == 0 else offs[1]
                    nb = 0
                    for dx, dy in ((0, 2 * ws), (2 * ws, 0), (-2 * ws, 0), (0, -2 * ws)):
                        nb += abs(int(pl[min(max(y0 + dy, 0), lh - 1), min(max(x0 + dx, 0), lw - 1)]) - cand)
                    bias += nb << ns
                out[z, wy, wx] = ((int(sad) << ds) + nw * bias) & 0xFFFFFFFF
    return out.astype(np.uint32)


def piecewise_offsets(rng, lh, lw, ws, amp):
    """Offsets constant on aligned windows of size 2*ws (the state at step 0 of a pass)."""
    p = 2 * ws
    base = rng.integers(-amp, amp + 1, (2, -(-lh // p), -(-lw // p)))
    return np.repeat(np.repeat(base, p, 1), p, 2)[:, :lh, :lw].astype(np.int16).copy()


@pytest.mark.parametrize("hdr", [False, True])
@pytest.mark.parametrize("W,H,S,rs,ws,R,iteration,step", [
    (64, 48, 64, 0, 8, 5, 0, 0), (64, 48, 80, 0, 4, 16, 5, 1), (66, 50, 66, 0, 2, 11, 6, 0), (128, 72, 128, 1, 16, 6, 4, 1),
    (130, 70, 192, 2, 2, 16, 7, 0), (96, 64, 96, 0, 32, 16, 1, 0),
])
def test_delta_sums_match_the_closed_form(hdr, W, H, S, rs, ws, R, iteration, step):
    rng = np.random.default_rng(W + ws + R)
    f1 = synth.make_random_frame(W, H, 11, hdr, S)
    f2 = synth.make_random_frame(W, H, 12, hdr, S)
    lw, lh = -(-W // (1 << rs)), -(-H // (1 << rs))
    offs = piecewise_offsets(rng, lh, lw, ws, 40)
    got = kernels.calc_delta_sums(f1, f2, offs, H, W, S, ws, R, rs, iteration, step, 8, 6, hdr)
    ref = closed_form_sums(f1, f2, offs, H, W, S, ws, R, rs, iteration, step, 8, 6, hdr)
    assert np.array_equal(got[:, ::ws, ::ws], ref)
    # everything that is not a window representative stays zero (atomics only hit representatives)
    mask = np.ones_like(got, bool)
    mask[:, ::ws, ::ws] = False
    assert not got[mask].any()


def test_delta_sums_wrap_modulo_2_32():
    """uint32 wrap-around is part of the contract: a 256x256 window of random bytes at deltaScalar 10."""
    W = H = 256
    f1 = synth.make_random_frame(W, H, 1)
    f2 = synth.make_random_frame(W, H, 2)
    offs = np.zeros((2, H, W), np.int16)
    got = kernels.calc_delta_sums(f1, f2, offs, H, W, W, 256, 5, 0, 0, 0, 10, 6, False)
    ref = closed_form_sums(f1, f2, offs, H, W, W, 256, 5, 0, 0, 0, 10, 6, False)
    assert np.array_equal(got[:, ::256, ::256], ref)
    Y1 = f1[:H * W].astype(np.int64)
    Y2 = f2[:H * W].astype(np.int64)
    assert (np.abs(Y1 - Y2).sum() << 10) > 2 ** 32  # the un-wrapped sum really exceeds 32 bits


def test_lowest_layer_ties_pick_the_lowest_index():
    sums = np.zeros((5, 4, 4), np.uint32)
    sums[:, 0, 0] = [7, 3, 3, 9, 3]
    sums[:, 0, 2] = [1, 1, 1, 1, 1]
    sums[:, 2, 0] = [9, 8, 7, 6, 5]
    sums[:, 2, 2] = [0xFFFFFFFF, 0xFFFFFFFE, 0xFFFFFFFF, 0xFFFFFFFE, 0xFFFFFFFF]
    layers = np.full((4, 4), 77, np.uint8)
    kernels.determine_lowest_layer(sums, layers, 2)
    assert layers[0, 0] == 1 and layers[0, 2] == 0 and layers[2, 0] == 4 and layers[2, 2] == 1
    assert layers[1, 1] == 77 and layers[0, 1] == 77  # non-representatives are untouched


def test_adjust_offset_array_adds_signed_squares():
    R = 16
    layers = np.zeros((8, 8), np.uint8)
    layers[0, 0], layers[0, 4], layers[4, 0], layers[4, 4] = 0, 8, 9, 15
    offs = np.full((2, 8, 8), 3, np.int16)
    kernels.adjust_offset_array(offs, layers, 4, R, 1)
    assert (offs[0] == 3).all()
    assert (offs[1, :4, :4] == 3 - 64).all() and (offs[1, :4, 4:] == 3).all() and (offs[1, 4:, :4] == 4).all() and (offs[1, 4:, 4:] == 3 + 49).all()


def test_blur_is_the_8x8_box_with_asymmetric_taps():
    rng = np.random.default_rng(5)
    offs = rng.integers(-300, 300, (2, 21, 37)).astype(np.int16)
    got = kernels.blur_flow(offs)

    def m(p, d):
        return d - (p - d + 1) if p >= d else (-p - 1 if p < 0 else p)

    for g, y, x in [(0, 0, 0), (1, 20, 36), (0, 10, 18), (1, 3, 35), (0, 19, 2)]:
        s = sum(int(offs[g, m(y + ky, 21), m(x + kx, 37)]) for ky in range(-4, 4) for kx in range(-4, 4))
        assert got[g, y, x] == int(s / 64)  # truncation toward zero
    assert (kernels.blur_flow(np.full((2, 9, 9), -7, np.int16)) == -7).all()


@pytest.mark.parametrize("hdr", [False, True])
def test_levels_identity_and_hdr_scale(hdr):
    """SDR levels 0/255 are the identity under IEEE division; HDR defaults scale by 65535/65280."""
    W, H = 256, 16
    dt = np.uint16 if hdr else np.uint8
    src = np.zeros((H + H // 2) * W, dt)
    vals = (np.arange(H * W) % 256).astype(dt)
    src[:H * W] = (vals.astype(np.uint16) << 8).astype(dt) if hdr else vals
    src[H * W:] = src[:H * W // 2]
    out = np.zeros_like(src)
    black, white = (0.0, 255.0 * 256.0) if hdr else (0.0, 255.0)
    for cz in (0, 1):
        kernels.copy_frame(src, out, H, W, W, W, black, white, cz, hdr)
    if not hdr:
        assert np.array_equal(out, src)
    else:
        y = src[:H * W].astype(np.float32)
        exp = np.minimum(y / np.float32(65280.0) * np.float32(65535.0), np.float32(65535.0)).astype(np.uint16)
        assert np.array_equal(out[:H * W], exp)


def test_ladder_window_sizes_match_the_survey():
    """ws0 and pass counts of SURVEY.md Appendix B."""
    for (W, H, maxres, hdr), (lw, lh, ws0, passes) in {
        (1920, 1080, 270, False): (480, 270, 256, 16), (1920, 1080, 540, False): (960, 540, 512, 18),
    }.items():
        o = OracleCalc(H, W, 0, 0, 8, 6, 0.0, 255.0, maxres, hdr)
        s = o.state()
        assert (s.flowWidth, s.flowHeight) == (lw, lh)
        o.close()
    o = OracleCalc(96, 160, 0, 0, 8, 6, 0.0, 255.0, 270, False)
    o.enableTaps()
    for t in range(3):
        o.updateFrame(synth.make_frame(160, 96, t))
    o.calculateOpticalFlow()
    assert o.numPasses() == 14 and o.passInfo(0)["windowSize"] == 128 and o.passInfo(13)["windowSize"] == 2
    # A.3 invariant: after the last pass the field is constant on 2x2 blocks
    off = o.readOffsetArray()
    assert np.array_equal(off[:, ::2, ::2].repeat(2, 1).repeat(2, 2)[:, :96, :160], off)


def test_flow_recovers_the_scene_motion_and_zero_for_identical_frames():
    W, H = 192, 128
    o = OracleCalc(H, W, 0, 0, 8, 6, 0.0, 255.0, 270, False)
    o.setParams(searchRadius=16)
    for t in range(3):
        o.updateFrame(synth.make_frame(W, H, t))
    o.calculateOpticalFlow()
    off = o.readOffsetArray()
    assert np.median(off[0]) == -synth.GLOBAL_MOTION[0] and np.median(off[1]) == -synth.GLOBAL_MOTION[1]
    f = synth.make_frame(W, H, 0)
    for _ in range(3):
        o.updateFrame(f)
    o.calculateOpticalFlow()
    assert not o.readOffsetArray().any() and not o.readFlow(latest=True).any()
    # two identical frames and zero flow: every blend is the frame itself (levels 0/255 are the identity)
    o.calculateOpticalFlow()
    for t in (0.0, 0.3, 1.0):
        o.warpFrames(t, 2)
        out = np.zeros_like(f)
        o.downloadFrame(out)
        Y, Yo = f[:H * W].reshape(H, W), out[:H * W].reshape(H, W)
        assert np.abs(Y[1:-1, 1:-1].astype(int) - Yo[1:-1, 1:-1].astype(int)).max() <= 1


def test_warp_mirror_never_samples_border_rows():
    """warp's mirror clamps to [1, dim-2]: mode 0 with zero flow reproduces the interior exactly."""
    W, H = 64, 32
    o = OracleCalc(H, W, 0, 0, 8, 6, 0.0, 255.0, 270, False)
    fs = [synth.make_random_frame(W, H, s) for s in (1, 2, 3)]
    for f in fs:
        o.updateFrame(f)
    o.warpFrames(0.5, 0)
    out = np.zeros_like(fs[0])
    o.downloadFrame(out)
    src, dst = fs[0][:H * W].reshape(H, W), out[:H * W].reshape(H, W)
    assert np.array_equal(dst[1:-1, 1:-1], src[1:-1, 1:-1])
    # row 0 -> 1 - 0 = 1 ; row H-1 -> (H-1) - 2*((H-1)-(H-2)) = H-3 (warpFrameKernelSDR.h:12-20)
    assert np.array_equal(dst[0, 1:-1], src[1, 1:-1]) and np.array_equal(dst[H - 1, 1:-1], src[H - 3, 1:-1])
    assert np.array_equal(dst[1:-1, 0], src[1:-1, 1]) and np.array_equal(dst[1:-1, W - 1], src[1:-1, W - 3])
    with pytest.raises(RuntimeError):
        o.warpFrames(1.0001, 2)
