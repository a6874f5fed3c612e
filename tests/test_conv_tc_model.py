"""CPU: a numpy model of csrc/conv_tc.cu's ADDRESSING -- pixel-tile decomposition, per-tap TMA box coordinates with
zero fill outside the image, the weight matrix's K ordering (tap-major, then channel) including the K blocks that
straddle a tap boundary when Cin % 64 != 0, and the epilogue's row -> pixel mapping and masks -- against F.conv2d.
The tensor-core / TMA / barrier mechanics are the flat GEMM's (validated on the GPU); what is new in the conv kernel
is exactly this index arithmetic, so it is pinned here where it can run without a GPU."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

BM, BK = 128, 64


def _box(x, c0, w0, h0, n, tw, th):
    """4-D TMA box {64 ch, tw, th, 1 image} at signed coordinates; elements outside the tensor read as zero.
    Returns the [th*tw, 64] tile in the order TMA writes it (w fastest, then h)."""
    N, H, W, C = x.shape
    out = np.zeros((th, tw, BK), np.float32)
    for i in range(th):
        for j in range(tw):
            h, w = h0 + i, w0 + j
            if 0 <= h < H and 0 <= w < W:
                c1 = min(c0 + BK, C)
                if c0 < C:
                    out[i, j, :c1 - c0] = x[n, h, w, c0:c1]
    return out.reshape(th * tw, BK)


def _wbox(wmat, k0, n0, BN):
    """2-D box {64 K, BN rows} of the [Cout, R*S*Cin] weight matrix, zero fill outside."""
    Cout, K = wmat.shape
    out = np.zeros((BN, BK), np.float32)
    r1, k1 = min(n0 + BN, Cout), min(k0 + BK, K)
    if n0 < Cout and k0 < K:
        out[:r1 - n0, :k1 - k0] = wmat[n0:r1, k0:k1]
    return out


def conv_model(x_nhwc, w_ocrs_c, ksize, dil, flat, in_affine=None):
    N, H, W, Cin = x_nhwc.shape
    Cout = w_ocrs_c.shape[0]
    wmat = w_ocrs_c.reshape(Cout, -1)                         # [Cout, R*S*Cin]: memory order of a channels-last weight
    BN = 256 if Cout > 128 else 128                           # host wrapper's channel-tile choice
    if flat:                                                  # host wrapper, ksize == 1
        xv = x_nhwc.reshape(1, 1, N * H * W, Cin)
        Nimg, Hh, Ww, log2_tw, tiles_h, tiles_w = 1, 1, N * H * W, 7, 1, (N * H * W + BM - 1) // BM
    else:
        xv = x_nhwc
        Nimg, Hh, Ww, log2_tw, tiles_h, tiles_w = N, H, W, 4, (H + 7) // 8, (W + 15) // 16
    tw, th = 1 << log2_tw, BM >> log2_tw
    out = np.full((Nimg * Hh * Ww, Cout), np.nan, np.float32)
    kb_per_tap = (Cin + BK - 1) // BK
    tiles_img = tiles_h * tiles_w
    for tm in range(Nimg * tiles_img):
        img, rem = divmod(tm, tiles_img)
        h0, w0 = (rem // tiles_w) * th, (rem % tiles_w) * tw
        for n0 in range(0, Cout, BN):
            acc = np.zeros((BM, BN), np.float32)
            order = [(tap, kb) for tap in range(ksize * ksize) for kb in range(kb_per_tap)]
            if in_affine is not None:                         # kXform: channel block outer, tap inner
                order = [(tap, kb) for kb in range(kb_per_tap) for tap in range(ksize * ksize)]
            for tap, kb in order:
                dh, dw = (tap // ksize - ksize // 2) * dil, (tap % ksize - ksize // 2) * dil
                a = _box(xv, kb * BK, w0 + dw, h0 + dh, img, tw, th)
                if in_affine is not None:                     # transform warps: per row (pixel) mask, per chunk channels
                    sc, sh = in_affine
                    for r in range(BM):
                        hh, ww = h0 + dh + (r >> log2_tw), w0 + dw + (r & (tw - 1))
                        if 0 <= hh < Hh and 0 <= ww < Ww:
                            for c in range(BK):
                                ch = kb * BK + c
                                s_, b_ = (sc[ch], sh[ch]) if ch < Cin else (0.0, 0.0)
                                a[r, c] = max(a[r, c] * s_ + b_, 0.0)
                b = _wbox(wmat, tap * Cin + kb * BK, n0, BN)
                acc += a @ b.T
            for m in range(BM):                               # epilogue: one thread per accumulator row
                h, w = h0 + (m >> log2_tw), w0 + (m & (tw - 1))
                if h < Hh and w < Ww:
                    pix = (img * Hh + h) * Ww + w
                    ncol = min(BN, Cout - n0)
                    assert np.isnan(out[pix, n0]), "pixel written twice"
                    out[pix, n0:n0 + ncol] = acc[m, :ncol]
    assert not np.isnan(out).any(), "pixel never written"
    return out.reshape(N, H, W, Cout)


@pytest.mark.parametrize("N,Cin,H,W,Cout,k,d", [(2, 72, 11, 19, 40, 3, 1), (1, 64, 20, 17, 136, 3, 3), (2, 8, 9, 9, 8, 3, 12),
                                                (3, 72, 5, 7, 24, 1, 1), (1, 128, 13, 11, 256, 1, 1)])
def test_conv_addressing_model(N, Cin, H, W, Cout, k, d):
    g = torch.Generator().manual_seed(N * 100 + Cin + d)
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g)
    ref = F.conv2d(x, w, None, 1, d * (k // 2), d).permute(0, 2, 3, 1).numpy()
    got = conv_model(x.permute(0, 2, 3, 1).contiguous().numpy(), w.permute(0, 2, 3, 1).contiguous().numpy(), k, d, flat=(k == 1))
    assert np.abs(got - ref).max() <= 2e-4 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("N,Cin,H,W,Cout,k,d", [(2, 72, 11, 19, 40, 3, 2), (2, 24, 5, 7, 16, 1, 1)])
def test_conv_in_transform_model(N, Cin, H, W, Cout, k, d):
    """kXform: relu(x * s + t) applied to the A tiles in shared memory; padding pixels and channels beyond Cin stay zero."""
    g = torch.Generator().manual_seed(Cin + k)
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g)
    sc, sh = torch.rand(Cin, generator=g) + 0.5, torch.randn(Cin, generator=g)
    ref = F.conv2d(F.relu(x * sc[None, :, None, None] + sh[None, :, None, None]), w, None, 1, d * (k // 2), d).permute(0, 2, 3, 1).numpy()
    got = conv_model(x.permute(0, 2, 3, 1).contiguous().numpy(), w.permute(0, 2, 3, 1).contiguous().numpy(), k, d, flat=(k == 1),
                     in_affine=(sc.numpy(), sh.numpy()))
    assert np.abs(got - ref).max() <= 2e-4 * max(1.0, np.abs(ref).max())


# ------------------------------------------------------------------ csrc/wgrad_tc.cu addressing
def _patch(t, c0, nch, w0, h0, n):
    """4-D TMA box {64 channels, 16 w, 4 h, 1 image} of an NHWC tensor at signed coordinates -> [64 pixels, nch]."""
    N, H, W, C = t.shape
    out = np.zeros((4, 16, nch), np.float32)
    for i in range(4):
        for j in range(16):
            h, w = h0 + i, w0 + j
            if 0 <= h < H and 0 <= w < W and c0 < C:
                c1 = min(c0 + nch, C)
                out[i, j, :c1 - c0] = t[n, h, w, c0:c1]
    return out.reshape(64, nch)


def wgrad_model(x_nhwc, g_nhwc, dil, splits):
    N, H, W, Cin = x_nhwc.shape
    Cout = g_nhwc.shape[3]
    tiles_h, tiles_w = (H + 3) // 4, (W + 15) // 16
    kb_total = N * tiles_h * tiles_w
    part = np.zeros((splits, 9, Cout, Cin), np.float32)
    for sp in range(splits):
        lo, hi = kb_total * sp // splits, kb_total * (sp + 1) // splits
        assert hi > lo
        for tap in range(9):
            dh, dw = (tap // 3 - 1) * dil, (tap % 3 - 1) * dil
            for co0 in range(0, Cout, 128):
                for ci0 in range(0, Cin, 256):
                    acc = np.zeros((128, 256), np.float32)
                    for kb in range(lo, hi):
                        img, rem = divmod(kb, tiles_h * tiles_w)
                        h0, w0 = (rem // tiles_w) * 4, (rem % tiles_w) * 16
                        a = _patch(g_nhwc, co0, 128, w0, h0, img)                 # [64 pix, 128 co]
                        b = _patch(x_nhwc, ci0, 256, w0 + dw, h0 + dh, img)       # [64 pix, 256 ci]
                        acc += a.T @ b
                    r1, c1 = min(128, Cout - co0), min(256, Cin - ci0)
                    part[sp, tap, co0:co0 + r1, ci0:ci0 + c1] = acc[:r1, :c1]
    return part.sum(0).reshape(3, 3, Cout, Cin).transpose(2, 3, 0, 1)


@pytest.mark.parametrize("N,Cin,H,W,Cout,d,splits", [(2, 24, 9, 19, 16, 1, 1), (1, 264, 7, 17, 136, 2, 3), (2, 8, 11, 5, 8, 5, 2)])
def test_wgrad_addressing_model(N, Cin, H, W, Cout, d, splits):
    g = torch.Generator().manual_seed(Cin + d)
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g, requires_grad=True)
    y = F.conv2d(x, w, None, 1, d, d)
    go = torch.randn(y.shape, generator=g)
    y.backward(go)
    got = wgrad_model(x.permute(0, 2, 3, 1).contiguous().numpy(), go.permute(0, 2, 3, 1).contiguous().numpy(), d, splits)
    assert np.abs(got - w.grad.numpy()).max() <= 2e-4 * max(1.0, float(w.grad.abs().max()))


# ------------------------------------------------------------------ csrc/pool.cu: scan order and tie rule vs ATen
def test_maxpool_tap_rule_matches_aten():
    """Forward kernel's rule -- windows scanned kh then kw over in-bounds taps, first tap initialises, a later tap wins only
    if strictly greater -- gives the same argmax as ATen (return_indices) on inputs full of ties."""
    g = torch.Generator().manual_seed(0)
    x = (F.relu(torch.randn(2, 3, 11, 14, generator=g)) * 4).round() / 4           # few distinct values: many ties
    out, idx = F.max_pool2d(x, 3, 2, 1, ceil_mode=True, return_indices=True)
    N, C, H, W = x.shape
    Ho, Wo = out.shape[2:]
    for n in range(N):
        for c in range(C):
            for ho in range(Ho):
                for wo in range(Wo):
                    best, tap = None, None
                    for kh in range(3):
                        h = 2 * ho - 1 + kh
                        if not 0 <= h < H:
                            continue
                        for kw in range(3):
                            w = 2 * wo - 1 + kw
                            if not 0 <= w < W:
                                continue
                            v = float(x[n, c, h, w])
                            if tap is None or v > best:
                                best, tap = v, (h, w)
                    assert best == float(out[n, c, ho, wo]) and tap[0] * W + tap[1] == int(idx[n, c, ho, wo])


def _im2col_tile(x_nhwc, m0, tap_r, tap_s, dil, pad, kb, rows=128, kch=64):
    """What ONE im2col-mode TMA load of conv_tc3.cu delivers (validated against the hardware by the GPU self-test): `rows`
    consecutive anchors of the bounding box (the image shifted by -pad), walked W first, then H, then N, starting at the
    tile's first pixel; every anchor reads pixel (anchor + tap offset) with offset {s*dil, r*dil}; channels [64 kb, 64 kb + 64);
    anything outside the tensor (padding, channel tail, anchors beyond the last image) is zero."""
    import numpy as np
    N, H, W, C = x_nhwc.shape
    out = np.zeros((rows, kch), np.float32)
    for i in range(rows):
        m = m0 + i
        n, rem = divmod(m, H * W)
        h, w = divmod(rem, W)
        if n >= N:
            continue                                                   # M tail: anchors past the last image
        hh, ww = (h - pad) + tap_r * dil, (w - pad) + tap_s * dil      # start coordinate {w0 - pad, h0 - pad} + instruction offsets
        if 0 <= hh < H and 0 <= ww < W:
            c0 = kb * kch
            c1 = min(C, c0 + kch)
            out[i, :c1 - c0] = x_nhwc[n, hh, ww, c0:c1]
    return out


@pytest.mark.parametrize("N,C,H,W,Co,k,d", [(2, 72, 9, 11, 40, 3, 2), (3, 64, 7, 5, 24, 3, 1), (1, 136, 13, 6, 8, 1, 1), (2, 64, 5, 5, 16, 3, 4)])
def test_flat_tile_im2col_addressing_model(N, C, H, W, Co, k, d):
    """conv_tc3.cu's producer loop as numpy: flat 128-pixel tiles (straddling rows and images), tap-outer / channel-block-inner
    K order with weight columns tap*Cin + 64 kb (a last block of a tap that runs into the next tap's columns meets zero-filled
    A channels), dilation, padding = dilation * (k // 2) -- summed over all K steps it must equal F.conv2d."""
    import numpy as np
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(N * 100 + C)
    x = rng.standard_normal((N, H, W, C)).astype(np.float32)
    wgt = rng.standard_normal((Co, k, k, C)).astype(np.float32)                        # [Cout, r, s, Cin] = the kernel's weight matrix rows
    wmat = wgt.reshape(Co, k * k * C)
    pad = d * (k // 2)
    M = N * H * W
    kb_per_tap = (C + 63) // 64
    got = np.zeros((M, Co), np.float32)
    for m0 in range(0, M, 128):
        acc = np.zeros((128, Co), np.float32)
        for tap in range(k * k):
            r, s = divmod(tap, k)
            for kb in range(kb_per_tap):
                a = _im2col_tile(x, m0, r, s, d, pad, kb) if k == 3 else _im2col_tile(x, m0, 0, 0, 1, 0, kb)
                col0 = tap * C + kb * 64
                b = np.zeros((Co, 64), np.float32)
                c1 = min(k * k * C, col0 + 64)
                b[:, :c1 - col0] = wmat[:, col0:c1]                                    # TMA zero-fills columns beyond K
                acc += a @ b.T
        rows = min(128, M - m0)
        got[m0:m0 + rows] = acc[:rows]
    ref = F.conv2d(torch.from_numpy(x).permute(0, 3, 1, 2), torch.from_numpy(wgt).permute(0, 3, 1, 2), None, 1, pad, d)
    ref = ref.permute(0, 2, 3, 1).reshape(M, Co).numpy()
    assert np.abs(got - ref).max() <= 1e-3 * max(1.0, np.abs(ref).max())
