// sgd_ema.cu -- the parameter update of one training step as ONE multi-tensor kernel:
//   SGD with momentum + weight decay (torch.optim.SGD semantics, built by lr_helper.py:12-27 / get_optimizer; the
//   per-group learning rate is whatever the poly schedule of lr_helper.py:78-113 wrote into the group), followed by the
//   EMA of the teacher parameters  t = d*t + (1-d)*s  (train_semi.py:531-548), in the same pass over the same values.
// Reference per step: ~360 parameter tensors x (weight-decay add, momentum mul/add, parameter add) + 360 x 3 kernels for
// the EMA -- about 1.5k tiny launches and ~2 GB of traffic.  Here: one launch, every byte moved once:
//   per element  read  p, g, m, t   (16 B)   write  p, m, t   (12 B)            [m = momentum buffer]
// The host passes a table of tensors (pointers, sizes, per-tensor lr / weight decay) and a table of chunks
// (tensor index, element offset); block b processes chunk b.  Arithmetic order follows torch.optim.SGD:
//   d = g + wd*p;  m = first ? d : momentum*m + d;  p = p - lr*m;  t = decay*t + (1-decay)*p  (rounded where torch rounds).
#include "common.cuh"

namespace u2pl {

struct SgdTensor {
    float *p; const float *g; float *m; float *t;       // t may be null (no EMA for this tensor)
    long long n;
    float lr, wd;
    int first;                                           // momentum buffer is uninitialised: m = d (torch: clone of d_p)
    int pad;
};

constexpr int kSgdChunk = 8192;                          // elements per block

__global__ void __launch_bounds__(256)
sgd_ema_kernel(const SgdTensor *__restrict__ tensors, const uint2 *__restrict__ chunks, float momentum, float decay, float one_minus, int do_ema)
{
    const uint2 ch = chunks[blockIdx.x];
    const SgdTensor T = tensors[ch.x];
    const long long off = static_cast<long long>(ch.y) * kSgdChunk;
    const int n = static_cast<int>(min(static_cast<long long>(kSgdChunk), T.n - off));
    float *p = T.p + off;
    const float *g = T.g + off;
    float *m = T.m + off;
    float *t = T.t ? T.t + off : nullptr;
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(t)) & 15) == 0;
    auto upd = [&](float &pv, float gv, float &mv, float &tv) {
        // rounding points of torch's multi-tensor SGD (_foreach_add(alpha) is an fma; _foreach_mul_ then _foreach_add_
        // round twice) and of the reference's EMA expression `d * t + (1 - d) * s` (three roundings)
        const float d = fmaf(T.wd, pv, gv);
        mv = T.first ? d : __fadd_rn(__fmul_rn(momentum, mv), d);
        pv = fmaf(-T.lr, mv, pv);
        if (do_ema) tv = __fadd_rn(__fmul_rn(decay, tv), __fmul_rn(one_minus, pv));
    };
    if (vec) {
        const int n4 = n >> 2;
        for (int i = threadIdx.x; i < n4; i += 256) {
            float4 pv = reinterpret_cast<float4 *>(p)[i];
            const float4 gv = reinterpret_cast<const float4 *>(g)[i];
            float4 mv = T.first ? make_float4(0, 0, 0, 0) : reinterpret_cast<float4 *>(m)[i];
            float4 tv = (t && do_ema) ? reinterpret_cast<float4 *>(t)[i] : make_float4(0, 0, 0, 0);
            upd(pv.x, gv.x, mv.x, tv.x); upd(pv.y, gv.y, mv.y, tv.y); upd(pv.z, gv.z, mv.z, tv.z); upd(pv.w, gv.w, mv.w, tv.w);
            reinterpret_cast<float4 *>(p)[i] = pv;
            reinterpret_cast<float4 *>(m)[i] = mv;
            if (t && do_ema) reinterpret_cast<float4 *>(t)[i] = tv;
        }
        for (int i = (n4 << 2) + threadIdx.x; i < n; i += 256) {
            float pv = p[i], mv = T.first ? 0.0f : m[i], tv = (t && do_ema) ? t[i] : 0.0f;
            upd(pv, g[i], mv, tv);
            p[i] = pv; m[i] = mv;
            if (t && do_ema) t[i] = tv;
        }
    } else {
        for (int i = threadIdx.x; i < n; i += 256) {
            float pv = p[i], mv = T.first ? 0.0f : m[i], tv = (t && do_ema) ? t[i] : 0.0f;
            upd(pv, g[i], mv, tv);
            p[i] = pv; m[i] = mv;
            if (t && do_ema) t[i] = tv;
        }
    }
}

}  // namespace u2pl

using namespace u2pl;

extern "C" int64_t u2pl_sgd_tensor_bytes(void) { return static_cast<int64_t>(sizeof(SgdTensor)); }
extern "C" int64_t u2pl_sgd_chunk_elems(void) { return kSgdChunk; }

extern "C" int u2pl_sgd_ema_step(const void *tensor_table, const void *chunk_table, int64_t n_chunks, float momentum, float ema_decay,
                                 float ema_one_minus, int do_ema, void *stream)
{
    if (!tensor_table || !chunk_table || n_chunks <= 0 || n_chunks >= (1LL << 31)) return bad_arg("sgd_ema_step: empty or oversized tables");
    sgd_ema_kernel<<<static_cast<unsigned>(n_chunks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const SgdTensor *>(tensor_table), static_cast<const uint2 *>(chunk_table), momentum, ema_decay, ema_one_minus, do_ema);
    return check_launch("sgd_ema_step");
}
