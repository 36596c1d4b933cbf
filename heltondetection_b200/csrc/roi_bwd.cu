// RoIAlign backward (what makes roi_align usable inside the reference's training loop, README.md:29).
// Restates torchvision's roi_align backward (ops/cpu/roi_align_kernel.cpp / cuda: bilinear_interpolate_gradient, one
// atomic add per corner): grad_input[b, c, y, x] += grad_output[k, c, ph, pw] * w_corner / count, sample by sample.
// NHWC grad_input (channels_last) is the fast layout: a thread owns a channel quad and adds it with ONE 128-bit
// reduction (red.global.add.v4.f32, sm_90+); the [C, PH*PW] grad_output tile of the RoI is staged in shared memory so
// the NCHW-ordered reads are coalesced.  Any other layout goes through the strided scalar-atomic kernel.
#include "hd_common.cuh"

struct RoiBwdParams {
    float* grad[HD_MAX_LEVELS];
    int H[HD_MAX_LEVELS], W[HD_MAX_LEVELS];
    float scale[HD_MAX_LEVELS];
    int n_levels, C, PH, PW, sampling_ratio, aligned;
    const float* rois; const int* level_ids; long long K;
    const float* grad_out;   // [K,C,PH,PW]
    long long sB, sC, sH, sW;   // element strides of grad[] (strided kernel)
};

struct RoiGeo { float sh, sw, bh, bw, count; int gh, gw, H, W, lvl, bidx; };

__device__ __forceinline__ RoiGeo roi_bwd_geo(const RoiBwdParams& p, long long k) {
    RoiGeo g;
    const float* roi = p.rois + k * 5;
    g.lvl = p.level_ids ? p.level_ids[k] : 0;
    g.H = p.H[g.lvl]; g.W = p.W[g.lvl];
    const float sc = p.scale[g.lvl];
    g.bidx = (int)roi[0];
    const float off = p.aligned ? 0.5f : 0.0f;
    g.sw = __fsub_rn(__fmul_rn(roi[1], sc), off); g.sh = __fsub_rn(__fmul_rn(roi[2], sc), off);
    const float ew = __fsub_rn(__fmul_rn(roi[3], sc), off), eh = __fsub_rn(__fmul_rn(roi[4], sc), off);
    float rw = __fsub_rn(ew, g.sw), rh = __fsub_rn(eh, g.sh);
    if (!p.aligned) { rw = fmaxf(rw, 1.0f); rh = fmaxf(rh, 1.0f); }
    g.bh = __fdiv_rn(rh, (float)p.PH); g.bw = __fdiv_rn(rw, (float)p.PW);
    g.gh = p.sampling_ratio > 0 ? p.sampling_ratio : (int)ceilf(__fdiv_rn(rh, (float)p.PH));
    g.gw = p.sampling_ratio > 0 ? p.sampling_ratio : (int)ceilf(__fdiv_rn(rw, (float)p.PW));
    g.count = (float)(g.gh * g.gw);
    return g;
}

// bilinear_interpolate_gradient of the reference: corner cells and weights of one sample; false if it contributes nothing
__device__ __forceinline__ bool roi_bwd_sample(float y, float x, int H, int W, int& yl, int& yh, int& xl, int& xh, float& w1, float& w2, float& w3,
                                               float& w4) {
    if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) return false;
    if (y <= 0.0f) y = 0.0f;
    if (x <= 0.0f) x = 0.0f;
    yl = (int)y; xl = (int)x;
    if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else yh = yl + 1;
    if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else xh = xl + 1;
    const float ly = __fsub_rn(y, (float)yl), lx = __fsub_rn(x, (float)xl), hy = __fsub_rn(1.0f, ly), hx = __fsub_rn(1.0f, lx);
    w1 = __fmul_rn(hy, hx); w2 = __fmul_rn(hy, lx); w3 = __fmul_rn(ly, hx); w4 = __fmul_rn(ly, lx);
    return true;
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ------------------------------------------------------------------------------------------------ NHWC, channel quads
__global__ void __launch_bounds__(256) roi_align_bwd_nhwc_quad_kernel(const __grid_constant__ RoiBwdParams p, int QT) {
    extern __shared__ __align__(16) float tile[];   // [C][PH*PW] grad_output of this RoI
    const long long k = blockIdx.x;
    const int nb = p.PH * p.PW, n = p.C * nb;
    const float* __restrict__ go = p.grad_out + (size_t)k * n;
    if ((((uintptr_t)go) & 15) == 0 && (n & 3) == 0) {
        for (int i = threadIdx.x; i < (n >> 2); i += blockDim.x) reinterpret_cast<float4*>(tile)[i] = hd_ldg_stream4(go + 4 * i);
    } else {
        for (int i = threadIdx.x; i < n; i += blockDim.x) tile[i] = hd_ldg_stream(go + i);
    }
    __syncthreads();
    const RoiGeo g = roi_bwd_geo(p, k);
    float* __restrict__ gi = p.grad[g.lvl] + (size_t)g.bidx * g.H * g.W * p.C;
    const int nq = p.C >> 2, groups = 256 / QT, grp = threadIdx.x / QT;
    for (int q = threadIdx.x % QT; q < nq; q += QT) {
        for (int bin = grp; bin < nb; bin += groups) {
            const int ph = bin / p.PW, pw = bin - ph * p.PW;
            const float* t = tile + (size_t)(4 * q) * nb + bin;
            const float g0 = t[0], g1 = t[nb], g2 = t[2 * nb], g3 = t[3 * nb];
            for (int iy = 0; iy < g.gh; ++iy) {
                const float y = __fadd_rn(__fadd_rn(g.sh, __fmul_rn((float)ph, g.bh)), __fdiv_rn(__fmul_rn(__fadd_rn((float)iy, 0.5f), g.bh), (float)g.gh));
                for (int ix = 0; ix < g.gw; ++ix) {
                    const float x = __fadd_rn(__fadd_rn(g.sw, __fmul_rn((float)pw, g.bw)), __fdiv_rn(__fmul_rn(__fadd_rn((float)ix, 0.5f), g.bw), (float)g.gw));
                    int yl, yh, xl, xh; float w1, w2, w3, w4;
                    if (!roi_bwd_sample(y, x, g.H, g.W, yl, yh, xl, xh, w1, w2, w3, w4)) continue;
                    float* c1 = gi + ((size_t)yl * g.W + xl) * p.C + 4 * q;
                    float* c2 = gi + ((size_t)yl * g.W + xh) * p.C + 4 * q;
                    float* c3 = gi + ((size_t)yh * g.W + xl) * p.C + 4 * q;
                    float* c4 = gi + ((size_t)yh * g.W + xh) * p.C + 4 * q;
#define RB(gv, w) __fdiv_rn(__fmul_rn(gv, w), g.count)
                    red_add_v4(c1, RB(g0, w1), RB(g1, w1), RB(g2, w1), RB(g3, w1));
                    red_add_v4(c2, RB(g0, w2), RB(g1, w2), RB(g2, w2), RB(g3, w2));
                    red_add_v4(c3, RB(g0, w3), RB(g1, w3), RB(g2, w3), RB(g3, w3));
                    red_add_v4(c4, RB(g0, w4), RB(g1, w4), RB(g2, w4), RB(g3, w4));
#undef RB
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ any layout (strides)
__global__ void __launch_bounds__(256) roi_align_bwd_strided_kernel(const __grid_constant__ RoiBwdParams p) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int nb = p.PH * p.PW;
    if (i >= p.K * p.C * nb) return;
    const int bin = (int)(i % nb);
    const int c = (int)((i / nb) % p.C);
    const long long k = i / ((long long)nb * p.C);
    const int ph = bin / p.PW, pw = bin - ph * p.PW;
    const RoiGeo g = roi_bwd_geo(p, k);
    const float gv = p.grad_out[i];
    float* __restrict__ gi = p.grad[g.lvl] + (size_t)g.bidx * p.sB + (size_t)c * p.sC;
    for (int iy = 0; iy < g.gh; ++iy) {
        const float y = __fadd_rn(__fadd_rn(g.sh, __fmul_rn((float)ph, g.bh)), __fdiv_rn(__fmul_rn(__fadd_rn((float)iy, 0.5f), g.bh), (float)g.gh));
        for (int ix = 0; ix < g.gw; ++ix) {
            const float x = __fadd_rn(__fadd_rn(g.sw, __fmul_rn((float)pw, g.bw)), __fdiv_rn(__fmul_rn(__fadd_rn((float)ix, 0.5f), g.bw), (float)g.gw));
            int yl, yh, xl, xh; float w1, w2, w3, w4;
            if (!roi_bwd_sample(y, x, g.H, g.W, yl, yh, xl, xh, w1, w2, w3, w4)) continue;
            atomicAdd(gi + (size_t)yl * p.sH + (size_t)xl * p.sW, __fdiv_rn(__fmul_rn(gv, w1), g.count));
            atomicAdd(gi + (size_t)yl * p.sH + (size_t)xh * p.sW, __fdiv_rn(__fmul_rn(gv, w2), g.count));
            atomicAdd(gi + (size_t)yh * p.sH + (size_t)xl * p.sW, __fdiv_rn(__fmul_rn(gv, w3), g.count));
            atomicAdd(gi + (size_t)yh * p.sH + (size_t)xh * p.sW, __fdiv_rn(__fmul_rn(gv, w4), g.count));
        }
    }
}

extern "C" HD_API int hd_roi_align_backward(const float* grad_out, const float* rois, const int32_t* level_ids, int64_t K, const hd_roi_level* levels,
                                            int n_levels, int layout, int C, int pooled_h, int pooled_w, int sampling_ratio, int aligned,
                                            void* stream) {
    HD_CHECK_ARG(levels != nullptr && n_levels >= 1 && n_levels <= HD_MAX_LEVELS, "n_levels must be in [1,%d], got %d", HD_MAX_LEVELS, n_levels);
    HD_CHECK_ARG(C >= 1 && pooled_h >= 1 && pooled_w >= 1, "bad C=%d or output size %dx%d", C, pooled_h, pooled_w);
    HD_CHECK_ARG(K >= 0, "K must be >= 0");
    HD_CHECK_ARG(layout == HD_LAYOUT_NCHW || layout == HD_LAYOUT_NHWC, "layout must be HD_LAYOUT_NCHW or HD_LAYOUT_NHWC, got %d", layout);
    HD_CHECK_ARG(n_levels == 1 || level_ids != nullptr || K == 0, "level_ids is NULL for a multi-level call");
    if (K == 0) return HD_OK;
    HD_CHECK_ARG(grad_out && rois, "null pointer");
    RoiBwdParams p;
    memset(&p, 0, sizeof(p));
    bool quad = (layout == HD_LAYOUT_NHWC) && (C % 4 == 0);
    for (int l = 0; l < n_levels; ++l) {
        HD_CHECK_ARG(levels[l].H > 0 && levels[l].W > 0 && levels[l].data != nullptr, "level %d: empty map or NULL grad buffer", l);
        p.grad[l] = const_cast<float*>(levels[l].data); p.H[l] = levels[l].H; p.W[l] = levels[l].W; p.scale[l] = levels[l].spatial_scale;
        quad = quad && (((uintptr_t)levels[l].data & 15) == 0);
    }
    p.n_levels = n_levels; p.C = C; p.PH = pooled_h; p.PW = pooled_w; p.sampling_ratio = sampling_ratio; p.aligned = aligned;
    p.rois = rois; p.level_ids = level_ids; p.K = K; p.grad_out = grad_out;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t tile = (size_t)C * pooled_h * pooled_w * 4;
    if (quad && tile <= 200 * 1024 && K < (1ll << 31)) {
        HD_ENSURE_SMEM(roi_align_bwd_nhwc_quad_kernel, 200 * 1024);
        int nq = C / 4, QT = 8;
        while (QT < nq && QT < 256) QT <<= 1;
        roi_align_bwd_nhwc_quad_kernel<<<(unsigned)K, 256, tile, st>>>(p, QT);
        HD_CUDA_LAUNCH_CHECK("roi_align_bwd_nhwc_quad_kernel");
        return HD_OK;
    }
    HD_CHECK_ARG(n_levels == 1, "multi-level RoIAlign backward needs the NHWC layout with C %% 4 == 0 and 16-byte aligned buffers");
    const long long H = levels[0].H, W = levels[0].W;
    if (layout == HD_LAYOUT_NHWC) { p.sB = H * W * C; p.sC = 1; p.sH = W * C; p.sW = C; }
    else { p.sB = (long long)C * H * W; p.sC = H * W; p.sH = W; p.sW = 1; }
    const long long total = K * C * pooled_h * pooled_w, blocks = (total + 255) / 256;
    HD_CHECK_ARG(blocks < (1ll << 31), "grid too large");
    roi_align_bwd_strided_kernel<<<(unsigned)blocks, 256, 0, st>>>(p);
    HD_CUDA_LAUNCH_CHECK("roi_align_bwd_strided_kernel");
    return HD_OK;
}

// ------------------------------------------------------------------------------------------------ RoIPool backward
// torchvision _roi_pool_backward: grad_input[b, c, argmax[k,c,ph,pw]] += grad_output[k,c,ph,pw]  (argmax = h*W + w, -1 = empty bin)
__global__ void __launch_bounds__(256) roi_pool_bwd_kernel(const float* __restrict__ grad_out, const int* __restrict__ argmax,
                                                           const float* __restrict__ rois, long long K, int C, int nb, float* __restrict__ gi,
                                                           long long sB, long long sC, long long sP) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K * C * nb) return;
    const int am = argmax[i];
    if (am < 0) return;
    const int c = (int)((i / nb) % C);
    const long long k = i / ((long long)nb * C);
    const int b = (int)rois[k * 5];
    atomicAdd(gi + (size_t)b * sB + (size_t)c * sC + (size_t)am * sP, grad_out[i]);
}

extern "C" HD_API int hd_roi_pool_backward(const float* grad_out, const int32_t* argmax, const float* rois, int64_t K, float* grad_in, int layout,
                                           int C, int H, int W, int pooled_h, int pooled_w, void* stream) {
    HD_CHECK_ARG(K >= 0 && C >= 1 && H >= 1 && W >= 1 && pooled_h >= 1 && pooled_w >= 1, "bad shape");
    HD_CHECK_ARG(layout == HD_LAYOUT_NCHW || layout == HD_LAYOUT_NHWC, "layout must be HD_LAYOUT_NCHW or HD_LAYOUT_NHWC, got %d", layout);
    if (K == 0) return HD_OK;
    HD_CHECK_ARG(grad_out && argmax && rois && grad_in, "null pointer");
    const int nb = pooled_h * pooled_w;
    const long long total = K * C * nb, blocks = (total + 255) / 256;
    HD_CHECK_ARG(blocks < (1ll << 31), "grid too large");
    const long long HW = (long long)H * W;
    const long long sB = HW * C, sC = (layout == HD_LAYOUT_NHWC) ? 1 : HW, sP = (layout == HD_LAYOUT_NHWC) ? C : 1;
    roi_pool_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(grad_out, argmax, rois, K, C, nb, grad_in, sB, sC, sP);
    HD_CUDA_LAUNCH_CHECK("roi_pool_bwd_kernel");
    return HD_OK;
}
