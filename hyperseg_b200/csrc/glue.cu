// Decoder glue around the patch kernels.
//
//  hsb_decoder_input_fwd   builds a level's input in one pass:
//        out = cat(coords, feature, bilinear_upsample(prev))          (reference hyperseg_v1_0.py:235-240:
//        F.interpolate(align_corners=False) + torch.cat + torch.cat = three kernels and two intermediate maps)
//  hsb_upsample_argmax_fwd the tail when only labels are wanted: bilinear upsample of the class logits to the
//        frame size (reference hyperseg_v1_0.py:250-251) fused with argmax over classes (hyperseg/test.py:171),
//        so the full-resolution logits never reach HBM.
//
// Both follow ATen's upsample_bilinear2d (align_corners=False): src = max((dst + 0.5) * in/out - 0.5, 0),
// i0 = floor(src), i1 = min(i0 + 1, in - 1), weights (1 - frac, frac), arithmetic in fp32, one rounding at the end.
#include "common.cuh"

namespace hsb {

struct BilinearAxis {
    int i0, i1;
    float w0, w1;
};

__device__ __forceinline__ BilinearAxis bilinear_axis(int dst, int in, float scale) {
    float src = fmaxf(((float)dst + 0.5f) * scale - 0.5f, 0.f);
    int i0 = (int)src;                        // src >= 0 -> truncation == floor
    i0 = min(i0, in - 1);
    BilinearAxis a;
    a.i0 = i0;
    a.i1 = min(i0 + 1, in - 1);
    a.w1 = src - (float)i0;
    a.w0 = 1.f - a.w1;
    return a;
}

struct DecInParams {
    const void* coords; const void* feat; const void* prev; void* out;
    int B, Cc, Cf, Cp, H, W, h, w;
    int64_t fsb, fsc, fsy, fsx;      // feature strides (elements): NCHW or NHWC
    float sy, sx;                    // h/H, w/W
};

// one thread = one output row segment of 8 pixels of one channel; channel order: coords | feature | prev
template <typename T>
__global__ void __launch_bounds__(256) decoder_input_kernel(const DecInParams p) {
    const int C = p.Cc + p.Cf + p.Cp;
    const int segs = (p.W + 7) / 8;
    const size_t total = (size_t)p.B * C * p.H * segs;
    const T* coords = reinterpret_cast<const T*>(p.coords);
    const T* feat = reinterpret_cast<const T*>(p.feat);
    const T* prev = reinterpret_cast<const T*>(p.prev);
    T* out = reinterpret_cast<T*>(p.out);
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int sg = idx % segs;
        const int y = (idx / segs) % p.H;
        const int c = (idx / ((size_t)segs * p.H)) % C;
        const int b = idx / ((size_t)segs * p.H * C);
        const int x0 = sg * 8;
        const int n = min(8, p.W - x0);
        float v[8];
        if (c < p.Cc) {
            const T* src = coords + ((size_t)c * p.H + y) * p.W + x0;
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = e < n ? ld_f(src + e) : 0.f;
        } else if (c < p.Cc + p.Cf) {
            const T* src = feat + (size_t)b * p.fsb + (size_t)(c - p.Cc) * p.fsc + (size_t)y * p.fsy + (size_t)x0 * p.fsx;
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = e < n ? ld_f(src + (size_t)e * p.fsx) : 0.f;
        } else {
            const int cp = c - p.Cc - p.Cf;
            const BilinearAxis ay = bilinear_axis(y, p.h, p.sy);
            const T* r0 = prev + (((size_t)b * p.Cp + cp) * p.h + ay.i0) * p.w;
            const T* r1 = prev + (((size_t)b * p.Cp + cp) * p.h + ay.i1) * p.w;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                if (e < n) {
                    const BilinearAxis ax = bilinear_axis(x0 + e, p.w, p.sx);
                    // ATen: h0lambda * (w0lambda * p00 + w1lambda * p01) + h1lambda * (w0lambda * p10 + w1lambda * p11)
                    v[e] = ay.w0 * (ax.w0 * ld_f(r0 + ax.i0) + ax.w1 * ld_f(r0 + ax.i1)) +
                           ay.w1 * (ax.w0 * ld_f(r1 + ax.i0) + ax.w1 * ld_f(r1 + ax.i1));
                } else {
                    v[e] = 0.f;
                }
            }
        }
        T* dst = out + (((size_t)b * C + c) * p.H + y) * p.W + x0;
        if (n == 8 && sizeof(T) == 2 && (p.W % 8) == 0) {
            __nv_bfloat162 q0 = __floats2bfloat162_rn(v[0], v[1]), q1 = __floats2bfloat162_rn(v[2], v[3]);
            __nv_bfloat162 q2 = __floats2bfloat162_rn(v[4], v[5]), q3 = __floats2bfloat162_rn(v[6], v[7]);
            uint4 pk = make_uint4(*reinterpret_cast<uint32_t*>(&q0), *reinterpret_cast<uint32_t*>(&q1),
                                  *reinterpret_cast<uint32_t*>(&q2), *reinterpret_cast<uint32_t*>(&q3));
            *reinterpret_cast<uint4*>(dst) = pk;
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e)
                if (e < n) st_f(dst + e, v[e]);
        }
    }
}

struct TailParams {
    const void* logits; unsigned char* labels;
    int B, C, h, w, H, W;
    float sy, sx;
};

// one thread = 4 consecutive output pixels; classes are walked with the four bilinear taps hoisted
template <typename T>
__global__ void __launch_bounds__(256) upsample_argmax_kernel(const TailParams p) {
    const int segs = (p.W + 3) / 4;
    const size_t total = (size_t)p.B * p.H * segs;
    const T* lg = reinterpret_cast<const T*>(p.logits);
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int sg = idx % segs;
        const int y = (idx / segs) % p.H;
        const int b = idx / ((size_t)segs * p.H);
        const int x0 = sg * 4;
        const BilinearAxis ay = bilinear_axis(y, p.h, p.sy);
        BilinearAxis ax[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) ax[e] = bilinear_axis(min(x0 + e, p.W - 1), p.w, p.sx);
        float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        int arg[4] = {0, 0, 0, 0};
        for (int c = 0; c < p.C; ++c) {
            const T* r0 = lg + (((size_t)b * p.C + c) * p.h + ay.i0) * p.w;
            const T* r1 = lg + (((size_t)b * p.C + c) * p.h + ay.i1) * p.w;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float v = ay.w0 * (ax[e].w0 * ld_f(r0 + ax[e].i0) + ax[e].w1 * ld_f(r0 + ax[e].i1)) +
                          ay.w1 * (ax[e].w0 * ld_f(r1 + ax[e].i0) + ax[e].w1 * ld_f(r1 + ax[e].i1));
                if (sizeof(T) == 2) v = __bfloat162float(__float2bfloat16_rn(v));   // the reference rounds the upsampled logits
                if (v > best[e]) { best[e] = v; arg[e] = c; }                        // first maximum wins, as torch.argmax
            }
        }
        unsigned char* dst = p.labels + ((size_t)b * p.H + y) * p.W + x0;
        if (x0 + 3 < p.W && (p.W % 4) == 0) {
            *reinterpret_cast<uchar4*>(dst) = make_uchar4((unsigned char)arg[0], (unsigned char)arg[1], (unsigned char)arg[2],
                                                          (unsigned char)arg[3]);
        } else {
            for (int e = 0; e < 4 && x0 + e < p.W; ++e) dst[e] = (unsigned char)arg[e];
        }
    }
}

}  // namespace hsb

using namespace hsb;

extern "C" int hsb_decoder_input_fwd(const void* coords, const void* feature, const void* prev, void* out,
                                     int B, int Cc, int Cf, int Cp, int H, int W, int h, int w,
                                     int64_t f_stride_b, int64_t f_stride_c, int64_t f_stride_y, int64_t f_stride_x,
                                     int dtype, void* stream) {
    HSB_REQUIRE(out && (Cc == 0 || coords) && (Cf == 0 || feature) && (Cp == 0 || prev), HSB_ERR_INVALID_ARG,
                "decoder_input: null pointer");
    HSB_REQUIRE(B > 0 && H > 0 && W > 0 && Cc >= 0 && Cf >= 0 && Cp >= 0 && Cc + Cf + Cp > 0, HSB_ERR_INVALID_ARG,
                "decoder_input: bad dimension");
    HSB_REQUIRE(Cp == 0 || (h > 0 && w > 0), HSB_ERR_INVALID_ARG, "decoder_input: bad source size");
    HSB_REQUIRE(dtype == HSB_F32 || dtype == HSB_BF16, HSB_ERR_INVALID_ARG, "decoder_input: bad dtype");
    DecInParams p;
    p.coords = coords; p.feat = feature; p.prev = prev; p.out = out;
    p.B = B; p.Cc = Cc; p.Cf = Cf; p.Cp = Cp; p.H = H; p.W = W; p.h = Cp ? h : 1; p.w = Cp ? w : 1;
    p.fsb = f_stride_b; p.fsc = f_stride_c; p.fsy = f_stride_y; p.fsx = f_stride_x;
    p.sy = (float)p.h / (float)H; p.sx = (float)p.w / (float)W;
    const size_t total = (size_t)B * (Cc + Cf + Cp) * H * ((W + 7) / 8);
    const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)std::max(1, device_sm_count()) * 32);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == HSB_F32) decoder_input_kernel<float><<<blocks, 256, 0, st>>>(p);
    else decoder_input_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(p);
    return check_launch("decoder_input launch");
}

extern "C" int hsb_upsample_argmax_fwd(const void* logits, void* labels, int B, int C, int h, int w, int H, int W,
                                       int dtype, void* stream) {
    HSB_REQUIRE(logits && labels, HSB_ERR_INVALID_ARG, "upsample_argmax: null pointer");
    HSB_REQUIRE(B > 0 && C > 0 && C <= 256 && h > 0 && w > 0 && H > 0 && W > 0, HSB_ERR_INVALID_ARG,
                "upsample_argmax: bad dimension (at most 256 classes fit a uint8 label)");
    HSB_REQUIRE(dtype == HSB_F32 || dtype == HSB_BF16, HSB_ERR_INVALID_ARG, "upsample_argmax: bad dtype");
    TailParams p;
    p.logits = logits; p.labels = (unsigned char*)labels; p.B = B; p.C = C; p.h = h; p.w = w; p.H = H; p.W = W;
    p.sy = (float)h / (float)H; p.sx = (float)w / (float)W;
    const size_t total = (size_t)B * H * ((W + 3) / 4);
    const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)std::max(1, device_sm_count()) * 32);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == HSB_F32) upsample_argmax_kernel<float><<<blocks, 256, 0, st>>>(p);
    else upsample_argmax_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(p);
    return check_launch("upsample_argmax launch");
}
