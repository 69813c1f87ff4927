// Patch-wise 1x1 convolution, bf16, persistent CTAs fed by a ring of TMA copies.
//
// Same operation as patch_conv1x1.cu (HyperPatchNoPadding.forward, reference hyperseg/models/hyperseg_v1_0.py:486-498, with
// the BatchNorm2d + ReLU of make_hyper_patch_conv2d_block :753-756 fused), for the layout the decoder produces: bf16,
// patch-major 16-byte aligned weight rows, groups == 1.  At the coarse decoder levels a patch is 1..16 pixels but 1.4..10 KB
// of weights, so the op is a stream of weight rows; this kernel keeps that stream in flight all the time:
//   * one CTA per SM walks a contiguous range of units (unit = PG neighbouring patches of one patch row);
//   * a unit's operands land in one of S shared-memory stages: the PG weight rows by cp.async.bulk, the x tile
//     [Cin][ph][PG * pw] by ONE 4-D TMA box, all byte-counted on the stage's "full" mbarrier; the producer (thread 0) refills
//     a stage as soon as every warp has released it ("empty" mbarrier), so S - 1 units are always in flight;
//   * every thread computes OB output channels x PB adjacent pixels in fp32 from packed bf16 pairs (weights along Cin,
//     pixels along W) and writes NCHW.  Lanes run over output channels where a unit has few pixels (weight rows of an odd
//     number of 32-bit words apart: conflict-free, x is a broadcast) and over pixels where it has many (x loads and y stores
//     are contiguous, weights are a broadcast).
// The one-shot kernel in patch_conv1x1.cu stays the general path (fp32, groups, NCHW / strided weights, odd shapes).
#include <cuda.h>
#include <mutex>

#include "bulk_copy.cuh"
#include "common.cuh"
#include "tcgen05.cuh"

namespace hsb {

void note_kernel(const char* name);

struct RingParams {
    const __nv_bfloat16* w;
    __nv_bfloat16* y;
    const float* post_scale;
    const float* post_shift;
    int act;
    int B, Cin, Cout, H, W, fh, fw, ph, pw;
    int PG, upr, units;          // patches per x box, x boxes per patch row, units in all
    int wsplit, PGw;             // a unit = one x box + the weight rows of PGw = PG / wsplit of its patches
    int hp, wrow;                // weights per patch, shared-memory pitch of a weight row (elements)
    int64_t w_row_stride;        // elements between the weight rows of neighbouring patches
    int stages, stage_bytes, x_off;      // ring depth, bytes per stage, offset of the x tile inside a stage
    int lanes_over_pixels;
};

constexpr int RING_MAX_THREADS = 512;
#ifndef HSB_RING_UNROLL
#define HSB_RING_UNROLL 8
#endif
constexpr int RING_UNROLL = HSB_RING_UNROLL;        // channel pairs per unrolled step of the inner loop

__device__ __forceinline__ float bf_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }

template <int OB, int PB>
__global__ void __launch_bounds__(RING_MAX_THREADS, 1) conv1x1_ring_kernel(const __grid_constant__ CUtensorMap xmap, const RingParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);           // [S]
    uint64_t* empty = full + 8;                                   // [S]
    unsigned char* stage0 = smem + 128;
    const int tid = threadIdx.x, lane = tid & 31;
    const int S = p.stages;

    const int u0 = (int)((long long)blockIdx.x * p.units / gridDim.x), u1 = (int)((long long)(blockIdx.x + 1) * p.units / gridDim.x);
    const int n = u1 - u0;
    if (tid == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, blockDim.x / 32); }
        mbar_fence_init();
        tma_prefetch_desc(&xmap);
    }
    __syncthreads();

    const int PGW = p.PG * p.pw, plane = p.ph * PGW;              // pixels per staged row / per channel
    const uint32_t w_bytes = (uint32_t)p.hp * 2, x_bytes = (uint32_t)(p.Cin * plane * 2);
    auto issue = [&](int k) {                                     // thread 0: operands of the CTA's k-th unit into stage k % S
        const int u = u0 + k, s = k % S;
        const int part = u % p.wsplit, box = u / p.wsplit, jg = box % p.upr, bi = box / p.upr, b = bi / p.fh, pi = bi % p.fh;
        unsigned char* st = stage0 + (size_t)s * p.stage_bytes;
        mbar_arrive_expect_tx(full + s, w_bytes * p.PGw + x_bytes);
        const __nv_bfloat16* wsrc = p.w + ((size_t)bi * p.fw + (size_t)jg * p.PG + (size_t)part * p.PGw) * p.w_row_stride;
        for (int q = 0; q < p.PGw; ++q)
            bulk_g2s(st + (size_t)q * p.wrow * 2, wsrc + (size_t)q * p.w_row_stride, w_bytes, full + s);
        tma_load_4d(st + p.x_off, &xmap, jg * PGW, pi * p.ph, 0, b, full + s);
    };
    if (tid == 0)
        for (int k = 0; k < min(n, S); ++k) issue(k);

    const int PWW = p.PGw * p.pw;                                 // pixels per row that belong to the unit's patches
    const int ogs = p.Cout / OB, pgs = p.ph * PWW / PB;           // output-channel groups, pixel groups per unit
    const int items = ogs * pgs;
    const size_t ostride = (size_t)p.H * p.W;
    for (int k = 0; k < n; ++k) {
        const int s = k % S, u = u0 + k;
        const int part = u % p.wsplit, box = u / p.wsplit, jg = box % p.upr, bi = box / p.upr, b = bi / p.fh, pi = bi % p.fh;
        const unsigned char* st = stage0 + (size_t)s * p.stage_bytes;
        const __nv_bfloat16* xs = reinterpret_cast<const __nv_bfloat16*>(st + p.x_off);
        mbar_wait(full + s, (k / S) & 1);
        for (int it = tid; it < items; it += blockDim.x) {
            int og, pg;
            if (p.lanes_over_pixels) { pg = it % pgs; og = it / pgs; } else { og = it % ogs; pg = it / ogs; }
            const int pl = pg * PB, r = pl / PWW, cl = pl % PWW, q = cl / p.pw;       // q: patch inside the unit's weight group
            const int col = part * PWW + cl, px = r * PGW + col;                      // column / pixel inside the x box
            // the thread's outputs are og, og + ogs, ...: neighbouring lanes read weight rows Cin / 2 words apart
            const __nv_bfloat16* wq = reinterpret_cast<const __nv_bfloat16*>(st) + (size_t)q * p.wrow + (size_t)og * p.Cin;
            const __nv_bfloat16* xc = xs + px;
            float acc[OB][PB];
#pragma unroll
            for (int o = 0; o < OB; ++o)
#pragma unroll
                for (int e = 0; e < PB; ++e) acc[o][e] = 0.f;
            const int wstep = ogs * p.Cin;
#pragma unroll RING_UNROLL
            for (int c = 0; c < p.Cin; c += 2) {
                float xa[PB], xb[PB];                             // channels c and c + 1 at the thread's pixels
                if (PB == 2) {
                    const uint32_t va = *reinterpret_cast<const uint32_t*>(xc + (size_t)c * plane);
                    const uint32_t vb = *reinterpret_cast<const uint32_t*>(xc + (size_t)(c + 1) * plane);
                    xa[0] = bf_lo(va); xa[PB - 1] = bf_hi(va); xb[0] = bf_lo(vb); xb[PB - 1] = bf_hi(vb);
                } else {
                    xa[0] = __uint_as_float((uint32_t) * reinterpret_cast<const unsigned short*>(xc + (size_t)c * plane) << 16);
                    xb[0] = __uint_as_float((uint32_t) * reinterpret_cast<const unsigned short*>(xc + (size_t)(c + 1) * plane) << 16);
                }
#pragma unroll
                for (int o = 0; o < OB; ++o) {
                    const uint32_t wv = *reinterpret_cast<const uint32_t*>(wq + (size_t)o * wstep + c);
                    const float w0 = bf_lo(wv), w1 = bf_hi(wv);
#pragma unroll
                    for (int e = 0; e < PB; ++e) acc[o][e] = fmaf(w1, xb[e], fmaf(w0, xa[e], acc[o][e]));
                }
            }
            __nv_bfloat16* yrow = p.y + (((size_t)b * p.Cout + og) * p.H + (size_t)pi * p.ph + r) * p.W + (size_t)jg * PGW + col;
#pragma unroll
            for (int o = 0; o < OB; ++o) {
                const int oc = og + o * ogs;
                float v[PB];
#pragma unroll
                for (int e = 0; e < PB; ++e) {
                    float t = acc[o][e];
                    if (p.post_scale) t = fmaf(t, __ldg(p.post_scale + oc), __ldg(p.post_shift + oc));
                    v[e] = act_apply(t, p.act);
                }
                __nv_bfloat16* dst = yrow + (size_t)o * ogs * ostride;
                if (PB == 2) *reinterpret_cast<__nv_bfloat162*>(dst) = __floats2bfloat162_rn(v[0], v[PB - 1]);
                else *dst = __float2bfloat16_rn(v[0]);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);                    // this warp has read everything it needs from the stage
        if (tid == 0 && k + S < n) {
            mbar_wait(empty + s, (k / S) & 1);                    // ... and so has every other warp
            issue(k + S);
        }
    }
}

typedef CUresult (*EncodeTiledFnR)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFnR ring_encode_fn() {
    static std::once_flag once;
    static EncodeTiledFnR fn = nullptr;
    std::call_once(once, [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFnR>(ptr);
        else
            cudaGetLastError();
    });
    return fn;
}

template <int OB, int PB>
static int launch_ring(const CUtensorMap& xmap, const RingParams& p, size_t smem, int grid, int threads, cudaStream_t st) {
    auto k = conv1x1_ring_kernel<OB, PB>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(HSB_ERR_CUDA, std::string("conv1x1_ring attr: ") + cudaGetErrorString(e));
    k<<<grid, threads, smem, st>>>(xmap, p);
    note_kernel("conv1x1_ring_kernel");
    return check_launch("conv1x1_ring launch");
}

// Runs the ring kernel when the problem fits it and sets *handled; otherwise leaves *handled false (the caller falls back
// to the general kernel).  Arguments as hsb_patch_conv1x1_fwd, bf16, patch-major weights.
int conv1x1_ring_try(const void* x, const void* w, void* y, const float* post_scale, const float* post_shift, int act, int B, int Cin,
                     int Cout, int H, int W, int fh, int fw, int groups, int64_t w_row_stride, cudaStream_t st, bool* handled) {
    *handled = false;
    static const bool off = [] { const char* v = getenv("HSB_NO_RING"); return v && v[0] == '1'; }();
    if (off || groups != 1 || (Cin & 1) || Cin > 256) return HSB_OK;
    const int ph = H / fh, pw = W / fw, hp = Cin * Cout;
    // Only one-pixel patches (the coarsest decoder level: 10 KB of weights per pixel) are a pure weight stream.  With 4 and
    // 16 pixels per patch the op is bound by fp32 FMA issue and load latency, and the one-shot kernel's 640 threads per SM
    // hide that better than one 256-thread CTA (measured at HyperSeg-M level 1 / 2: 12.6 / 10.5 us against 14.1 / 10.2 us
    // here; HSB_RING_ALL=1 forces this kernel for experiments).
    static const bool all = [] { const char* v = getenv("HSB_RING_ALL"); return v && v[0] == '1'; }();
    if (ph * pw > 1 && !all) return HSB_OK;
    if (ph > 256 || ((uintptr_t)x & 15) || ((uintptr_t)w & 15) || ((uintptr_t)y & 3) || (W * 2) % 16 || (hp * 2) % 16 || (w_row_stride * 2) % 16) return HSB_OK;
    // patches per unit: the x box must have 16-byte rows (PG * pw * 2 bytes) and PG must divide fw; the smallest such PG that
    // gives every thread an item (units stay small: more units than SMs, several stages in flight), else the largest one
    const int PB = (pw % 2 == 0) ? 2 : 1;
    static const int env_threads = [] { const char* v = getenv("HSB_RING_THREADS"); return v ? atoi(v) : 0; }();
    static const int env_items = [] { const char* v = getenv("HSB_RING_ITEMS"); return v ? atoi(v) : 0; }();
    const int threads = env_threads >= 64 && env_threads <= RING_MAX_THREADS && env_threads % 32 == 0 ? env_threads : 256;
    const int min_items = env_items > 0 ? env_items : threads;
    const int wrow_e = ((Cin * Cout + 7) / 8 * 8) + ((((Cin * Cout + 7) / 8) % 2 == 0) ? 8 : 0);
    // x box: the smallest PG with 16-byte rows.  Its patches are split over `wsplit` units (each loads the whole, small x box
    // and its own PGw weight rows): units stay small -- several per SM for balance, 3-4 stages in flight -- as long as a unit
    // still has an item for every thread with one output channel per thread.
    int PG = 0;
    for (int cand = 1; cand <= fw && cand <= 32; ++cand)
        if (fw % cand == 0 && (cand * pw * 2) % 16 == 0 && cand * pw <= 256) { PG = cand; break; }
    if (!PG) return HSB_OK;
    const int PGW = PG * pw, plane = ph * PGW;
    // (measured: splitting further, with one output channel per thread, is slower -- the kernel is bound by instruction
    // latency, not by balance -- so a unit is split only while two output channels per thread still fill the CTA)
    int wsplit = 1;
    while (PG % (wsplit * 2) == 0 && ((PG / (wsplit * 2)) * pw) % PB == 0 && Cout % 2 == 0 &&
           (Cout / 2) * (ph * (PG / (wsplit * 2)) * pw / PB) >= min_items) wsplit *= 2;
    const int PGw = PG / wsplit, unit_pg = ph * PGw * pw / PB;
    const int OB = (Cout % 4 == 0 && Cout / 4 * unit_pg >= min_items) ? 4 : (Cout % 2 == 0 && Cout / 2 * unit_pg >= min_items) ? 2 : 1;
    RingParams p;
    p.w = (const __nv_bfloat16*)w; p.y = (__nv_bfloat16*)y; p.post_scale = post_scale; p.post_shift = post_shift; p.act = act;
    p.B = B; p.Cin = Cin; p.Cout = Cout; p.H = H; p.W = W; p.fh = fh; p.fw = fw; p.ph = ph; p.pw = pw;
    p.PG = PG; p.upr = fw / PG; p.wsplit = wsplit; p.PGw = PGw; p.units = B * fh * p.upr * wsplit; p.hp = hp; p.w_row_stride = w_row_stride;
    p.wrow = wrow_e;              // an odd number of 16-byte units: the rows of neighbouring patches start in different banks
    p.x_off = (int)(((size_t)PGw * p.wrow * 2 + 127) / 128 * 128);
    p.stage_bytes = (int)(((size_t)p.x_off + (size_t)Cin * plane * 2 + 127) / 128 * 128);
    p.stages = std::min(4, (int)((200 * 1024) / p.stage_bytes));
    if (p.stages < 2) return HSB_OK;
    p.lanes_over_pixels = unit_pg >= 32;
    const size_t smem = 128 + (size_t)p.stages * p.stage_bytes;
    EncodeTiledFnR encode = ring_encode_fn();
    if (!encode) return HSB_OK;
    CUtensorMap xmap;
    const cuuint64_t dim[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)Cin, (cuuint64_t)B};
    const cuuint64_t str[3] = {(cuuint64_t)W * 2, (cuuint64_t)H * W * 2, (cuuint64_t)Cin * H * W * 2};
    const cuuint32_t box[4] = {(cuuint32_t)PGW, (cuuint32_t)ph, (cuuint32_t)Cin, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    if (encode(&xmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dim, str, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return HSB_OK;
    const int grid = std::min(p.units, std::max(1, device_sm_count()));
    static const bool verbose = [] { const char* v = getenv("HSB_VERBOSE"); return v && v[0] == '1'; }();
    if (verbose)
        fprintf(stderr, "[hsb] conv1x1_ring: %d -> %d, %dx%d px/patch: x box %d patches, %d weight rows per unit, %d units, OB %d PB %d, %d stages of %d B, %d threads\n",
                Cin, Cout, ph, pw, PG, PGw, p.units, OB, PB, p.stages, p.stage_bytes, threads);
    *handled = true;
    if (PB == 2)
        return OB == 4 ? launch_ring<4, 2>(xmap, p, smem, grid, threads, st)
             : OB == 2 ? launch_ring<2, 2>(xmap, p, smem, grid, threads, st) : launch_ring<1, 2>(xmap, p, smem, grid, threads, st);
    return OB == 4 ? launch_ring<4, 1>(xmap, p, smem, grid, threads, st)
         : OB == 2 ? launch_ring<2, 1>(xmap, p, smem, grid, threads, st) : launch_ring<1, 1>(xmap, p, smem, grid, threads, st);
}

}  // namespace hsb
