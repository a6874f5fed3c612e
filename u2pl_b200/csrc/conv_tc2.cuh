// conv_tc2.cuh -- interface between conv_tc.cu (C-ABI entry points, 1-CTA kernels) and conv_tc2.cu (CTA-pair kernel).
#pragma once
#include <cstdint>

namespace u2pl {

struct Conv2Params {
    int Nimg, H, W, Cin, Cout;        // logical NHWC geometry the tensor maps were built from (flat 1x1: Nimg=H=1, W=N*H*W)
    int R, S, dil;
    int log2_tw, tiles_h, tiles_w;    // 128-pixel tile = TH x TW, TW = 1 << log2_tw; tiles per image
    const float *scale, *shift;       // [Cout] or null
    int has_residual, relu;
    float *stat_part;                 // kStats: [2 * pixel tiles][2][Cout]
};

bool conv_tc2_eligible(int64_t cout, bool xform);
int64_t conv_tc2_stat_parts(int64_t n, int64_t h, int64_t w, int ksize);
int conv_tc2_launch(const void *x, const void *wgt, void *out, int64_t n, int64_t h, int64_t w, int64_t cin, int64_t cout,
                    int ksize, int dilation, const float *scale, const float *shift, const void *residual, int relu,
                    float *stat_part, const char *what, void *stream);

}  // namespace u2pl
