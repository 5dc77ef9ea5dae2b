"""Host logic behind the flat-position kernels (csrc/train_tc5.cu, train_wgrad5.cu, train_gemm.cu): positions of the pitch-(W+1)
pixel array are decoded with multiply-high division by ceil(2^32 / P).  The launchers only take that path for positions below
2^32 / P; this checks the arithmetic the guard relies on, and the position -> pixel map itself against a plain loop."""
import numpy as np
import pytest


@pytest.mark.parametrize('P', [4, 7, 9, 13, 17, 25, 33, 49, 65, 97, 111, 129])
def test_multiply_high_quotient_is_exact_below_2_32_over_p(P):
    magic = ((1 << 32) + P - 1) // P
    lim = (1 << 32) // P
    rs = np.random.RandomState(P)
    g = np.concatenate([np.arange(0, 100000, dtype=np.uint64), rs.randint(0, lim, 500000).astype(np.uint64),
                        np.arange(lim - 100000, lim, dtype=np.uint64)])
    q = (g * np.uint64(magic)) >> np.uint64(32)
    assert np.array_equal(q, g // np.uint64(P))


@pytest.mark.parametrize('geom', [(2, 3, 5), (3, 6, 8), (1, 12, 16)])
def test_flat_position_to_pixel_map(geom):
    """Position g of the flat array (pitch P = W + 1, H + 1 rows per image, column 0 / row 0 = padding) <-> NHWC pixel, and
    a tap (dy, dx) is the constant shift (dy - 1) P + (dx - 1): the padded neighbour of every pixel is a padding slot."""
    N, H, W = geom
    P, RPI = W + 1, H + 1
    magicP, magicR = ((1 << 32) + P - 1) // P, ((1 << 32) + RPI - 1) // RPI

    def pixel(flat):                       # what w5_pixel / t5f_pixel compute (shifted by one image block)
        gs = flat + RPI * P
        if gs < 0:
            return -1
        R = (gs * magicP) >> 32
        X = gs - R * P
        nn = (R * magicR) >> 32
        yy = R - nn * RPI
        n = nn - 1
        if n < 0 or n >= N or X == 0 or yy == 0:
            return -1
        return (n * H + yy - 1) * W + X - 1

    seen = set()
    for flat in range(-P - 1, N * RPI * P + P + 2):
        px = pixel(flat)
        if px < 0:
            continue
        assert px not in seen
        seen.add(px)
        n, rem = divmod(px, H * W)
        y, x = divmod(rem, W)
        assert flat == (n * RPI + y + 1) * P + x + 1
        for dy in range(3):
            for dx in range(3):
                q = pixel(flat + (dy - 1) * P + (dx - 1))
                yy, xx = y + dy - 1, x + dx - 1
                want = (n * H + yy) * W + xx if 0 <= yy < H and 0 <= xx < W else -1
                assert q == want, (flat, dy, dx, q, want)
    assert len(seen) == N * H * W
