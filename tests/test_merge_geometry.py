"""CPU checks of the integer geometry the class-major merge kernels rely on (vfmseg_b200/csrc/slide_tail.cuh:
slide_merge_class_kernel, ms_merge_class_kernel). The kernels replace the float source-index expression of PyTorch's
upsample_bilinear2d(align_corners=False) — the one mmseg's resize() runs inside slide_inference
(Ms_VFM_encoder_decoder.py:455-461) and ms_inference (:449-459) — by shifts and masks for power-of-two upsampling factors,
stage window footprints with replicated borders, and pick taps by strip phase. This file restates that index logic in numpy
float32 / Python ints and compares it, pixel by pixel, with the direct expression: same taps, same weights (bit-equal
float32), every tap inside the staged footprint, for the tile / thread mapping the kernels use."""
import numpy as np
import pytest

f32 = np.float32
TW, TH, FR, FC, MAXW = 64, 16, 6, 20, 4   # MERGE_TW / MERGE_TH / MERGE_FR / MERGE_FC / MERGE_MAXW


def _src(c, inv):
    v = f32(inv) * (f32(c) + f32(0.5)) - f32(0.5)
    return f32(0) if v < 0 else v


def _direct(o, inv, n):
    """bilerp_setup along one axis: (first tap, second tap, weight of the second tap)."""
    s = _src(o, inv)
    i0 = min(int(s), n - 1)
    return i0, i0 + (1 if i0 < n - 1 else 0), float(f32(s) - f32(i0))


def _boxes(H, W, crop, stride):
    from vfmseg_b200.engine import slide_boxes
    return slide_boxes(H, W, (crop, crop), (stride, stride))


def _strip_from_footprint(lg, ih, iw, r_lo, r_hi, c_lo, c_hi, row, c0, limit, rowok):
    """One thread's strip of one source as the kernels compute it: per pixel (row tap 0, row tap 1, column tap 0, column tap 1,
    h1, w1) or None, plus the in-window mask."""
    S, inv = 1 << lg, 1.0 / (1 << lg)
    ly_lo, lx_lo = int(_src(r_lo, inv)), int(_src(c_lo, inv))
    fr, fc = int(_src(r_hi, inv)) + 2 - ly_lo, int(_src(c_hi, inv)) + 2 - lx_lo
    assert fr <= FR and fc <= 18 and fr * fc <= 128           # one staging pass of 128 threads per class parity
    res, inmask = [None] * 4, 0
    anyin = rowok and c0 + 3 >= 0 and c0 < limit
    if not anyin:
        return res, inmask
    fast = c0 >= S // 2 and c0 + 3 < limit
    sy = _src(row, inv)
    y0 = int(sy)
    h1 = float(f32(sy) - f32(y0))
    rr = y0 - ly_lo

    def staged(r, c):                                          # replicated borders: the source index is clamped while staging
        assert 0 <= r < fr and 0 <= c < fc, (r, c, fr, fc)
        return min(ly_lo + r, ih - 1), min(lx_lo + c, iw - 1)

    for j in range(4):
        if rowok and 0 <= c0 + j < limit:
            inmask |= 1 << j
    if fast:
        t0 = c0 - S // 2
        base, ph0 = t0 >> lg, t0 & (S - 1)
        kx = min(S - ph0, 4)                                   # first pixel one source column further
        w1_0 = f32((f32(ph0) + f32(0.5)) * f32(inv))
        assert base - lx_lo + 2 < FC                           # the third column load stays inside the slot row
        for j in range(4):
            o = 1 if j >= kx else 0
            w1 = float(f32(f32(w1_0 + f32(j) * f32(inv)) - f32(o)))
            col = base - lx_lo + o
            tl, tr, bl = staged(rr, col), staged(rr, col + 1), staged(rr + 1, col)
            staged(rr + 1, col + 1)
            res[j] = (tl[0], bl[0], tl[1], tr[1], h1, w1)
    else:
        for j in range(4):
            c = c0 + j
            if c < 0 or c >= limit:
                continue
            sx = _src(c, inv)
            x0 = int(sx)
            col = x0 - lx_lo
            tl, tr, bl = staged(rr, col), staged(rr, col + 1), staged(rr + 1, col)
            res[j] = (tl[0], bl[0], tl[1], tr[1], h1, float(f32(sx) - f32(x0)))
    return res, inmask


def _threads(tx0, ty0, H, W):
    for t in range(256):
        lane = t & 31
        xs, y = tx0 + ((t >> 5) * 2 + (lane >> 4)) * 4, ty0 + (lane & 15)   # warp = 2 strips x 16 rows
        if xs < W and y < H:
            yield xs, y


@pytest.mark.parametrize("H,W,crop,stride,lg0,lgr", [(96, 160, 64, 40, 3, 4), (64, 128, 64, 43, 2, 2), (96, 132, 64, 21, 2, 5),
                                                     (128, 192, 128, 85, 3, 4)])
def test_class_major_merge_geometry_equals_bilinear_source_index(H, W, crop, stride, lg0, lgr):
    lh, lw, rh = H >> lg0, W >> lg0, crop >> lgr
    boxes = _boxes(H, W, crop, stride)
    checked = 0
    for ty0 in range(0, H, TH):
        for tx0 in range(0, W, TW):
            over = [(by, bx) for (by, bx) in boxes
                    if max(ty0 - by, 0) <= min(ty0 + TH - 1 - by, crop - 1) and max(tx0 - bx, 0) <= min(tx0 + TW - 1 - bx, crop - 1)]
            for xs, y in _threads(tx0, ty0, H, W):
                # context source (ms_merge_class_kernel): the tile in image coordinates, factor 2^lg0
                res, _ = _strip_from_footprint(lg0, lh, lw, ty0, min(ty0 + TH - 1, H - 1), tx0, min(tx0 + TW - 1, W - 1), y, xs, W, True)
                for j in range(4):
                    ya, yb, h1 = _direct(y, 1.0 / (1 << lg0), lh)
                    xa, xb, w1 = _direct(xs + j, 1.0 / (1 << lg0), lw)
                    assert res[j] == (ya, yb, xa, xb, h1, w1)
                    checked += 1
                # window sources: factor 2^lgr (lgr = 2 is slide_merge_class_kernel's x4 LinearHead logits)
                for by, bx in over[:MAXW]:
                    cy, cx0 = y - by, xs - bx
                    rowok = 0 <= cy < crop
                    res, inmask = _strip_from_footprint(lgr, rh, rh, max(ty0 - by, 0), min(ty0 + TH - 1 - by, crop - 1), max(tx0 - bx, 0),
                                                        min(tx0 + TW - 1 - bx, crop - 1), cy, cx0, crop, rowok)
                    for j in range(4):
                        inside = rowok and 0 <= cx0 + j < crop
                        assert bool(inmask & (1 << j)) == inside
                        if inside:
                            ya, yb, h1 = _direct(cy, 1.0 / (1 << lgr), rh)
                            xa, xb, w1 = _direct(cx0 + j, 1.0 / (1 << lgr), rh)
                            assert res[j] == (ya, yb, xa, xb, h1, w1)
                            checked += 1
    assert checked > 10000


def test_x4_strip_phase_is_uniform_per_window_and_weights_are_the_four_immediates():
    """slide_merge_class_kernel: for exact x4 upsampling a strip (xs % 4 == 0) of a window at column bx has phase (xs - bx - 2) & 3
    whatever xs, and the horizontal weights are 0.125 / 0.375 / 0.625 / 0.875 exactly."""
    for bx in range(0, 23):
        phases = {(xs - bx - 2) & 3 for xs in range(((bx + 2 + 3) // 4) * 4, 512, 4)}
        assert len(phases) == 1
    for cx in range(2, 600):
        x0, _, w1 = _direct(cx, 0.25, 1 << 20)
        assert x0 == (cx - 2) >> 2 and w1 == 0.125 + 0.25 * ((cx - 2) & 3) and float(f32(1) - f32(w1)) == 1.0 - w1
    for cx in (0, 1):
        assert _direct(cx, 0.25, 128) == (0, 1, 0.0)


def test_full_size_tiles_need_at_most_four_window_footprints():
    """BASELINE configs 2 / 3 / 4 (1024x2048, crop 512, stride 341 / 320) and the BDD100K test size (1024x1820): no 64x16 tile lies
    under more than MERGE_MAXW windows, so the per-pixel gather inside the class loop never runs there."""
    for H, W, crop, stride in [(1024, 2048, 512, 341), (1024, 2048, 512, 320), (1024, 1820, 512, 341), (1024, 2048, 1024, 682)]:
        boxes = _boxes(H, W, crop, stride)
        worst = 0
        for ty0 in range(0, H, TH):
            for tx0 in range(0, W, TW):
                n = sum(1 for (by, bx) in boxes
                        if max(ty0 - by, 0) <= min(ty0 + TH - 1 - by, crop - 1) and max(tx0 - bx, 0) <= min(tx0 + TW - 1 - bx, crop - 1))
                worst = max(worst, n)
        assert 1 <= worst <= MAXW, (H, W, crop, stride, worst)
