"""Host-side restatement of the index logic of the strip x2 resampling kernels (csrc/elementwise.cu,
upsample2_fwd_strip_kernel / upsample2_bwd_strip_kernel): the invariants the kernels are sized by, checked in the
float arithmetic the kernels use (ATen upsample_bilinear2d source coordinates, reference call sites
sepconv/model.py:213-234 nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True) and the F.interpolate calls
of the other backbones).  No GPU: the CUDA kernels themselves are held to torch's upsample in test_kernels_gpu.py."""
import random

import numpy as np

F32 = np.float32
UP_STRIP_MAX, UP_SLOTS, UP_CAND = 8, 6, 6          # constants of elementwise.cu


def up2_scale(n):
    return F32(n - 1) / F32(2 * n - 1) if 2 * n > 1 else F32(0)


def up2_src(o, n, align, scale):
    """up2_src of elementwise.cu: source rows i0, i1 and the weight t of i1 for output index o."""
    if align:
        s = F32(scale) * F32(o)
    else:
        s = F32(0.5) * (F32(o) + F32(0.5)) - F32(0.5)
        if s < 0:
            s = F32(0)
    i0 = min(int(s), n - 1)
    i1 = i0 + (1 if i0 < n - 1 else 0)
    return i0, i1, F32(s - F32(i0))


def test_contributing_outputs_lie_in_the_candidate_window():
    """Input index g only receives from outputs 2g-2 .. 2g+3 (UP_CAND = 6 columns, the row walk starts at 2g-2)."""
    lo, hi = 99, -99
    for n in list(range(1, 400)) + [512, 1024, 2048, 4096]:
        for align in (0, 1):
            sc = up2_scale(n)
            for o in range(2 * n):
                a, b, t = up2_src(o, n, align, sc)
                for g, w in ((a, 1 - t), (b, t)):
                    if w != 0:
                        lo, hi = min(lo, o - 2 * g), max(hi, o - 2 * g)
    assert -2 <= lo and hi <= 3, (lo, hi)
    assert hi - lo + 1 <= UP_CAND


def test_a_strip_of_output_rows_touches_at_most_up_slots_source_rows():
    for n in list(range(1, 300)) + [1000]:
        for align in (0, 1):
            sc = up2_scale(n)
            for strip in (1, 2, 4, 8):
                for oy0 in range(0, 2 * n, strip):
                    rows = [up2_src(o, n, align, sc)[:2] for o in range(oy0, min(oy0 + strip, 2 * n))]
                    assert rows[-1][1] - rows[0][0] + 1 <= UP_SLOTS
                    assert rows[-1][1] - rows[0][0] + 1 <= strip // 2 + 2


def _reference_bwd(r, full_h, align, ly0, h, hy0, oh):
    dx = np.zeros(h)
    sc = up2_scale(full_h)
    for o in range(hy0, hy0 + oh):
        a, b, t = up2_src(o, full_h, align, sc)
        for row, w in ((a, 1 - t), (b, t)):
            if 0 <= row - ly0 < h:
                dx[row - ly0] += w * r[o - hy0]
    return dx


def _strip_walk_bwd(r, full_h, align, ly0, h, hy0, oh, strip):
    """The walk of upsample2_bwd_strip_kernel over one column: two running accumulators, rows finished in order."""
    dx = np.full(h, np.nan)
    sc = up2_scale(full_h)
    for yy0 in range(0, h, strip):
        yy1 = min(yy0 + strip, h) - 1
        gy0, gy1 = yy0 + ly0, yy1 + ly0
        cur, acc0, acc1 = gy0, 0.0, 0.0

        def finish(g, a):
            assert np.isnan(dx[g - ly0]), "a row is stored exactly once"
            dx[g - ly0] = a
        o_lo, o_hi = max(hy0, 2 * gy0 - 2), min(hy0 + oh - 1, 2 * gy1 + 3)
        assert o_hi - o_lo + 1 <= 2 * UP_STRIP_MAX + UP_CAND
        for o in range(o_lo, o_hi + 1):
            a, b, t = up2_src(o, full_h, align, sc)
            if b < gy0 or a > gy1:
                continue
            while cur < a:
                finish(cur, acc0)
                acc0, acc1, cur = acc1, 0.0, cur + 1
            rr = r[o - hy0]
            if a == cur:
                acc0 += (1 - t) * rr
                if b == a:
                    acc0 += t * rr
                else:
                    acc1 += t * rr
            else:
                assert b == cur and a == cur - 1, (a, b, cur)
                acc0 += t * rr
        while cur <= gy1:
            finish(cur, acc0)
            acc0, acc1, cur = acc1, 0.0, cur + 1
    return dx


def test_strip_walk_equals_the_scatter_form_on_full_grids_and_windows():
    rnd = random.Random(1)
    rng = np.random.default_rng(2)
    for _ in range(1500):
        full_h = rnd.randint(1, 40)
        align = rnd.random() < 0.6
        strip = rnd.choice((1, 2, 4, 8))
        if rnd.random() < 0.4:
            ly0, h, hy0, oh = 0, full_h, 0, 2 * full_h
        else:   # a window of outputs and a low-resolution window that covers their sources (as ops.py requires)
            hy0 = rnd.randint(0, 2 * full_h - 1)
            oh = rnd.randint(1, 2 * full_h - hy0)
            sc = up2_scale(full_h)
            srcs = [up2_src(o, full_h, align, sc) for o in range(hy0, hy0 + oh)]
            lo, hi = min(s[0] for s in srcs), max(s[1] for s in srcs)
            ly0 = rnd.randint(max(0, lo - 3), lo)
            h = rnd.randint(hi - ly0 + 1, min(full_h - ly0, hi - ly0 + 4))
        r = rng.random(oh)
        want = _reference_bwd(r, full_h, align, ly0, h, hy0, oh)
        got = _strip_walk_bwd(r, full_h, align, ly0, h, hy0, oh, strip)
        assert np.allclose(want, got, atol=1e-12), (full_h, align, ly0, h, hy0, oh, strip)
