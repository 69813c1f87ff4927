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
    int out_channels;                // channel count of `out` (== Cc + Cf + Cp unless the upsampled part is done elsewhere)
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
        T* dst = out + (((size_t)b * p.out_channels + c) * p.H + y) * p.W + x0;
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

// ---- exact 2x upsampling (every decoder level and the logit tail): fixed taps, vector loads ----------------------------------
// For out = 2*in, align_corners=False:  out[2j] = 0.25*in[j-1] + 0.75*in[j],  out[2j+1] = 0.75*in[j] + 0.25*in[j+1]  (indices
// clamped to the border; at the border the two taps coincide, which reproduces ATen's src = max(src, 0) / i1 = min(i0+1, in-1)).
// A thread takes source rows (i, i+1) and 4 source columns c0..c0+3 (+ one neighbour each side) and produces the 2 x 8 output
// pixels (rows 2i+1, 2i+2; columns 2*c0 .. 2*c0+7) that depend on them.

__device__ __forceinline__ void load_row6(const __nv_bfloat16* row, int c0, int w, float (&s)[6]) {
    // s[0..5] = in[c0-1 .. c0+4], clamped; c0 % 4 == 0 and w % 4 == 0 -> the middle four are one aligned 8-byte load
    const uint2 mid = *reinterpret_cast<const uint2*>(row + c0);
    const __nv_bfloat162 m0 = *reinterpret_cast<const __nv_bfloat162*>(&mid.x), m1 = *reinterpret_cast<const __nv_bfloat162*>(&mid.y);
    s[1] = __low2float(m0); s[2] = __high2float(m0); s[3] = __low2float(m1); s[4] = __high2float(m1);
    s[0] = c0 > 0 ? __bfloat162float(row[c0 - 1]) : s[1];
    s[5] = c0 + 4 < w ? __bfloat162float(row[c0 + 4]) : s[4];
}

__device__ __forceinline__ void hinterp8(const float (&s)[6], float (&o)[8]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        o[2 * j] = 0.25f * s[j] + 0.75f * s[j + 1];           // ATen: w0*p0 + w1*p1 with (w0, w1) = (0.25, 0.75)
        o[2 * j + 1] = 0.75f * s[j + 1] + 0.25f * s[j + 2];
    }
}

__global__ void __launch_bounds__(256) upsample2x_argmax_kernel(const TailParams p) {
    const int segs = p.w / 4;
    const size_t total = (size_t)p.B * (p.h + 1) * segs;
    const __nv_bfloat16* lg = reinterpret_cast<const __nv_bfloat16*>(p.logits);
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int sg = idx % segs;
        const int i = (int)((idx / segs) % (p.h + 1)) - 1;          // source row pair (i, i+1), i = -1 .. h-1
        const int b = idx / ((size_t)segs * (p.h + 1));
        const int r0 = max(i, 0), r1 = min(i + 1, p.h - 1), c0 = sg * 4;
        float best1[8], best2[8];
        int arg1[8], arg2[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { best1[e] = best2[e] = -INFINITY; arg1[e] = arg2[e] = 0; }
        // (unrolling this loop 4x / 8x was measured: 57.8 / 42.1 us against 43.5 us -- the kernel is bound by the ~100 fp32
        // instructions per channel and thread, not by load latency)
        for (int c = 0; c < p.C; ++c) {
            const __nv_bfloat16* base = lg + ((size_t)b * p.C + c) * p.h * p.w;
            float s0[6], s1[6], h0[8], h1[8];
            load_row6(base + (size_t)r0 * p.w, c0, p.w, s0);
            load_row6(base + (size_t)r1 * p.w, c0, p.w, s1);
            hinterp8(s0, h0);
            hinterp8(s1, h1);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                // the reference rounds the upsampled logits to bf16 before argmax
                const float v1 = __bfloat162float(__float2bfloat16_rn(0.75f * h0[e] + 0.25f * h1[e]));   // row 2i+1
                const float v2 = __bfloat162float(__float2bfloat16_rn(0.25f * h0[e] + 0.75f * h1[e]));   // row 2i+2
                if (v1 > best1[e]) { best1[e] = v1; arg1[e] = c; }
                if (v2 > best2[e]) { best2[e] = v2; arg2[e] = c; }
            }
        }
        const int y1 = 2 * i + 1, y2 = 2 * i + 2;
        if (y1 >= 0) {
            uint2 pk;
            pk.x = arg1[0] | (arg1[1] << 8) | (arg1[2] << 16) | (arg1[3] << 24);
            pk.y = arg1[4] | (arg1[5] << 8) | (arg1[6] << 16) | (arg1[7] << 24);
            *reinterpret_cast<uint2*>(p.labels + ((size_t)b * p.H + y1) * p.W + 2 * c0) = pk;
        }
        if (y2 < p.H) {
            uint2 pk;
            pk.x = arg2[0] | (arg2[1] << 8) | (arg2[2] << 16) | (arg2[3] << 24);
            pk.y = arg2[4] | (arg2[5] << 8) | (arg2[6] << 16) | (arg2[7] << 24);
            *reinterpret_cast<uint2*>(p.labels + ((size_t)b * p.H + y2) * p.W + 2 * c0) = pk;
        }
    }
}

// upsampled channels of the level input (bf16, exact 2x): out[b, c_off + cp, 2i+1 / 2i+2, 8 px]
__global__ void __launch_bounds__(256) decoder_prev2x_kernel(const DecInParams p) {
    const int segs = p.w / 4;
    const size_t total = (size_t)p.B * p.Cp * (p.h + 1) * segs;
    const int C = p.Cc + p.Cf + p.Cp;
    const __nv_bfloat16* prev = reinterpret_cast<const __nv_bfloat16*>(p.prev);
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out);
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int sg = idx % segs;
        const int i = (int)((idx / segs) % (p.h + 1)) - 1;
        const int cp = (idx / ((size_t)segs * (p.h + 1))) % p.Cp;
        const int b = idx / ((size_t)segs * (p.h + 1) * p.Cp);
        const int r0 = max(i, 0), r1 = min(i + 1, p.h - 1), c0 = sg * 4;
        const __nv_bfloat16* base = prev + ((size_t)b * p.Cp + cp) * p.h * p.w;
        float s0[6], s1[6], h0[8], h1[8];
        load_row6(base + (size_t)r0 * p.w, c0, p.w, s0);
        load_row6(base + (size_t)r1 * p.w, c0, p.w, s1);
        hinterp8(s0, h0);
        hinterp8(s1, h1);
        __nv_bfloat16* o = out + (((size_t)b * C + p.Cc + p.Cf + cp) * p.H) * p.W + 2 * c0;
        const int y1 = 2 * i + 1, y2 = 2 * i + 2;
        if (y1 >= 0) {
            uint32_t q[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                __nv_bfloat162 t = __floats2bfloat162_rn(0.75f * h0[2 * e] + 0.25f * h1[2 * e], 0.75f * h0[2 * e + 1] + 0.25f * h1[2 * e + 1]);
                q[e] = *reinterpret_cast<uint32_t*>(&t);
            }
            *reinterpret_cast<uint4*>(o + (size_t)y1 * p.W) = make_uint4(q[0], q[1], q[2], q[3]);
        }
        if (y2 < p.H) {
            uint32_t q[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                __nv_bfloat162 t = __floats2bfloat162_rn(0.25f * h0[2 * e] + 0.75f * h1[2 * e], 0.25f * h0[2 * e + 1] + 0.75f * h1[2 * e + 1]);
                q[e] = *reinterpret_cast<uint32_t*>(&t);
            }
            *reinterpret_cast<uint4*>(o + (size_t)y2 * p.W) = make_uint4(q[0], q[1], q[2], q[3]);
        }
    }
}

// feature channels of the level input when the encoder hands them over channels-last (bf16): one thread = 8 pixels x 8
// channels.  It reads eight 16-byte pixel rows (neighbouring lanes take the neighbouring channel groups of the same pixels, so a
// warp consumes whole contiguous runs), transposes the 8 x 8 tile in registers and writes eight 16-byte channel rows.  The
// generic kernel above reads one channel per thread: at 16 channels every 32-byte sector was fetched 16 times (70 us for the
// last level of HyperSeg-M, as long as the fused MetaBlock that consumes it).
__global__ void __launch_bounds__(256) decoder_feat_nhwc_kernel(const DecInParams p) {
    const int segs = p.W / 8, cgs = p.Cf / 8;
    const size_t total = (size_t)p.B * p.H * segs * cgs;
    const __nv_bfloat16* feat = reinterpret_cast<const __nv_bfloat16*>(p.feat);
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out);
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int cg = idx % cgs;
        const int sg = (idx / cgs) % segs;
        const int y = (idx / ((size_t)cgs * segs)) % p.H;
        const int b = idx / ((size_t)cgs * segs * p.H);
        const __nv_bfloat16* src = feat + (size_t)b * p.fsb + (size_t)y * p.fsy + (size_t)(sg * 8) * p.fsx + cg * 8;
        uint4 in[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) in[e] = __ldg(reinterpret_cast<const uint4*>(src + (size_t)e * p.fsx));
        __nv_bfloat16* dst = out + (((size_t)b * p.out_channels + p.Cc + cg * 8) * p.H + y) * p.W + sg * 8;
        const size_t plane = (size_t)p.H * p.W;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            // channel c of pixels (2k, 2k + 1): low or high half of word c / 2 of both pixel rows
            const uint32_t sel = (c & 1) ? 0x7632u : 0x5410u;
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t a = reinterpret_cast<const uint32_t*>(&in[2 * k])[c >> 1];
                const uint32_t bq = reinterpret_cast<const uint32_t*>(&in[2 * k + 1])[c >> 1];
                w[k] = __byte_perm(a, bq, sel);
            }
            *reinterpret_cast<uint4*>(dst + (size_t)c * plane) = make_uint4(w[0], w[1], w[2], w[3]);
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
    p.out_channels = Cc + Cf + Cp;
    p.fsb = f_stride_b; p.fsc = f_stride_c; p.fsy = f_stride_y; p.fsx = f_stride_x;
    p.sy = (float)p.h / (float)H; p.sx = (float)p.w / (float)W;
    cudaStream_t st = (cudaStream_t)stream;
    const int cap = std::max(1, device_sm_count()) * 32;
    // exact 2x upsampling of bf16 maps (every level of the shipped decoders): fixed-tap vector kernel for the
    // upsampled channels, the generic kernel then only copies coords + feature
    const bool fast2x = dtype == HSB_BF16 && Cp > 0 && H == 2 * p.h && W == 2 * p.w && (p.w % 4) == 0 &&
                        ((uintptr_t)prev % 16) == 0 && ((uintptr_t)out % 16) == 0;
    if (fast2x) {
        const size_t t2 = (size_t)B * Cp * (p.h + 1) * (p.w / 4);
        decoder_prev2x_kernel<<<(int)std::min<size_t>((t2 + 255) / 256, cap), 256, 0, st>>>(p);
        int rc = check_launch("decoder_prev2x launch");
        if (rc != HSB_OK) return rc;
        if (Cc + Cf == 0) return HSB_OK;
        DecInParams q = p;
        q.Cp = 0;                       // copy channels only; the output keeps its full channel count via out_channels
        q.out_channels = Cc + Cf + Cp;
        // channels-last feature with whole 8-channel groups: transposing copy; the generic kernel then only writes the coords
        const char* no_nhwc = getenv("HSB_NO_NHWC_GLUE");      // A/B switch for scripts/time_glue.py
        const bool nhwc8 = !(no_nhwc && no_nhwc[0] == '1') && Cf > 0 && Cf % 8 == 0 && W % 8 == 0 && f_stride_c == 1 && f_stride_x % 8 == 0 && f_stride_y % 8 == 0 &&
                           f_stride_b % 8 == 0 && ((uintptr_t)feature % 16) == 0;
        if (nhwc8) {
            const size_t tf = (size_t)B * H * (W / 8) * (Cf / 8);
            decoder_feat_nhwc_kernel<<<(int)std::min<size_t>((tf + 255) / 256, cap), 256, 0, st>>>(q);
            rc = check_launch("decoder_feat_nhwc launch");
            if (rc != HSB_OK || Cc == 0) return rc;
            q.Cf = 0;                   // coords only ...
            q.out_channels = Cc + Cf + Cp;
        }
        const size_t t1 = (size_t)B * (q.Cc + q.Cf) * H * ((W + 7) / 8);
        decoder_input_kernel<__nv_bfloat16><<<(int)std::min<size_t>((t1 + 255) / 256, cap), 256, 0, st>>>(q);
        return check_launch("decoder_input launch");
    }
    const size_t total = (size_t)B * (Cc + Cf + Cp) * H * ((W + 7) / 8);
    const int blocks = (int)std::min<size_t>((total + 255) / 256, cap);
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
    cudaStream_t st = (cudaStream_t)stream;
    const int cap = std::max(1, device_sm_count()) * 32;
    if (dtype == HSB_BF16 && H == 2 * h && W == 2 * w && (w % 4) == 0 && ((uintptr_t)logits % 8) == 0 && ((uintptr_t)labels % 8) == 0) {
        const size_t t2 = (size_t)B * (h + 1) * (w / 4);
        upsample2x_argmax_kernel<<<(int)std::min<size_t>((t2 + 255) / 256, cap), 256, 0, st>>>(p);
        return check_launch("upsample2x_argmax launch");
    }
    const size_t total = (size_t)B * H * ((W + 3) / 4);
    const int blocks = (int)std::min<size_t>((total + 255) / 256, cap);
    if (dtype == HSB_F32) upsample_argmax_kernel<float><<<blocks, 256, 0, st>>>(p);
    else upsample_argmax_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(p);
    return check_launch("upsample_argmax launch");
}
