// conv_tc3.cuh -- interface between conv_tc.cu (C-ABI entry points) and conv_tc3.cu (flat-tile im2col-TMA kernel).
#pragma once
#include <cstdint>

namespace u2pl {

int64_t conv_tc3_stat_parts(int64_t n, int64_t h, int64_t w);
int conv_tc3_launch(const void *x, const void *wgt, void *out, int64_t n, int64_t h, int64_t w, int64_t cin, int64_t cout,
                    int ksize, int dilation, const float *scale, const float *shift, const void *residual, int relu,
                    float *stat_part, const char *what, void *stream);

}  // namespace u2pl
