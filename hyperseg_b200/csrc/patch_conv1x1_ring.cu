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

// ---- the same ring, arithmetic on the tensor cores (mma.sync m16n8k16, bf16 x bf16 -> fp32) -------------------------------
// One warp owns one patch of the unit at a time: D[Cout x pixels] = W[Cout x Cin] . X[Cin x pixels].  The weight row is read
// from shared memory exactly once, as A fragments (32-bit loads of channel pairs: no unpacking, no FMA instructions -- a
// one-pixel patch needs ~160 warp instructions instead of ~500); X comes as B fragments (the patch's pixels are the N
// columns: 8 per tile, zero-filled past the patch).  MT = Cout tiles of 16, NTP = pixel tiles of 8 per patch.
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

constexpr int MMA_CONSUMERS = 8, MMA_THREADS = 32 * (MMA_CONSUMERS + 1), MMA_MAX_SLOTS = 16, MMA_MAX_STAGES = 4;
constexpr int MMA_BAR_BYTES = 2048;       // full_w / empty_w [stage][slot], full_x / empty_x [stage]

template <int MT, int NTP>
__global__ void __launch_bounds__(MMA_THREADS, 1) conv1x1_mma_kernel(const __grid_constant__ CUtensorMap xmap, const RingParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    // The ring is refilled per PATCH SLOT, not per unit: a warp releases the weight row of its patch as soon as it has
    // consumed it and the producer warp streams the row of the unit S steps ahead into that slot at once, so the S x PG rows
    // of the ring are in flight all the time (with whole-unit stages only one 84 KB copy was in flight after the first two).
    uint64_t* full_w = reinterpret_cast<uint64_t*>(smem);                       // [S][MMA_MAX_SLOTS]
    uint64_t* empty_w = full_w + MMA_MAX_STAGES * MMA_MAX_SLOTS;                // [S][MMA_MAX_SLOTS]
    uint64_t* full_x = empty_w + MMA_MAX_STAGES * MMA_MAX_SLOTS;                // [S]
    uint64_t* empty_x = full_x + MMA_MAX_STAGES;                                // [S]
    unsigned char* stage0 = smem + MMA_BAR_BYTES;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;                    // fragment coordinates: group id, thread in group
    const int S = p.stages;
    const int u0 = (int)((long long)blockIdx.x * p.units / gridDim.x), u1 = (int)((long long)(blockIdx.x + 1) * p.units / gridDim.x);
    const int n = u1 - u0;
    const int users = min(MMA_CONSUMERS, p.PG);               // warps that own at least one slot: they release the x tile
    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            for (int q = 0; q < p.PG; ++q) { mbar_init(full_w + s * MMA_MAX_SLOTS + q, 1); mbar_init(empty_w + s * MMA_MAX_SLOTS + q, 1); }
            mbar_init(full_x + s, 1);
            mbar_init(empty_x + s, users);
        }
        mbar_fence_init();
        tma_prefetch_desc(&xmap);
    }
    __syncthreads();
    const int PGW = p.PG * p.pw, plane = p.ph * PGW, npix = p.ph * p.pw;
    const uint32_t w_bytes = (uint32_t)p.hp * 2, x_bytes = (uint32_t)(p.Cin * plane * 2);

    if (warp == MMA_CONSUMERS) {
        // ================= producer warp (one lane) =================
        if (lane == 0) {
            for (int k = 0; k < n; ++k) {
                const int u = u0 + k, s = k % S;
                const uint32_t par = ((k / S) & 1) ^ 1;               // the slot's previous user (unit k - S) has released it
                const int jg = u % p.upr, bi = u / p.upr, b = bi / p.fh, pi = bi % p.fh;
                unsigned char* st = stage0 + (size_t)s * p.stage_bytes;
                if (k >= S) mbar_wait(empty_x + s, par);
                mbar_arrive_expect_tx(full_x + s, x_bytes);
                tma_load_4d(st + p.x_off, &xmap, jg * PGW, pi * p.ph, 0, b, full_x + s);
                const __nv_bfloat16* wsrc = p.w + ((size_t)bi * p.fw + (size_t)jg * p.PG) * p.w_row_stride;
                for (int q = 0; q < p.PG; ++q) {
                    if (k >= S) mbar_wait(empty_w + s * MMA_MAX_SLOTS + q, par);
                    mbar_arrive_expect_tx(full_w + s * MMA_MAX_SLOTS + q, w_bytes);
                    bulk_g2s(st + (size_t)q * p.wrow * 2, wsrc + (size_t)q * p.w_row_stride, w_bytes, full_w + s * MMA_MAX_SLOTS + q);
                }
            }
        }
        return;
    }

    // per lane: offsets of its B-fragment pixels inside the x tile (column n = g of every pixel tile), -1 past the patch
    int xoff[NTP];
#pragma unroll
    for (int j = 0; j < NTP; ++j) {
        const int pp = j * 8 + g;
        xoff[j] = pp < npix ? (pp / p.pw) * PGW + pp % p.pw : -1;
    }
    // epilogue constants of the lane's output channels (rows g and g + 8 of every Cout tile)
    float sc[MT][2], sh[MT][2];
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int oc = m * 16 + h * 8 + g;
            sc[m][h] = (p.post_scale && oc < p.Cout) ? __ldg(p.post_scale + oc) : 1.f;
            sh[m][h] = (p.post_scale && oc < p.Cout) ? __ldg(p.post_shift + oc) : 0.f;
        }
    const int ksteps = (p.Cin + 15) / 16;
    const size_t ostride = (size_t)p.H * p.W;
    for (int k = 0; k < n; ++k) {
        const int s = k % S, u = u0 + k;
        const uint32_t par = (k / S) & 1;
        const int jg = u % p.upr, bi = u / p.upr, b = bi / p.fh, pi = bi % p.fh;
        const unsigned char* st = stage0 + (size_t)s * p.stage_bytes;
        const unsigned short* xs = reinterpret_cast<const unsigned short*>(st + p.x_off);
        if (warp < users) mbar_wait(full_x + s, par);
        for (int q = warp; q < p.PG; q += MMA_CONSUMERS) {
            mbar_wait(full_w + s * MMA_MAX_SLOTS + q, par);
            const __nv_bfloat16* wq = reinterpret_cast<const __nv_bfloat16*>(st) + (size_t)q * p.wrow;
            const unsigned short* xq = xs + q * p.pw;
            float acc[MT][NTP][4];
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int j = 0; j < NTP; ++j)
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[m][j][e] = 0.f;
#pragma unroll 2
            for (int ks = 0; ks < ksteps; ++ks) {
                const int k0 = ks * 16 + 2 * t;                               // this lane's channel pairs: k0, k0 + 8
                const bool v0 = k0 < p.Cin, v1 = k0 + 8 < p.Cin;              // Cin is even: a pair is valid or not as a whole
                uint32_t bf[NTP][2];
#pragma unroll
                for (int j = 0; j < NTP; ++j) {
                    bf[j][0] = bf[j][1] = 0u;
                    if (xoff[j] >= 0) {
                        const unsigned short* xp = xq + xoff[j];
                        if (v0) bf[j][0] = (uint32_t)xp[(size_t)k0 * plane] | ((uint32_t)xp[(size_t)(k0 + 1) * plane] << 16);
                        if (v1) bf[j][1] = (uint32_t)xp[(size_t)(k0 + 8) * plane] | ((uint32_t)xp[(size_t)(k0 + 9) * plane] << 16);
                    }
                }
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                    const int r0 = m * 16 + g, r1 = r0 + 8;
                    const uint32_t* w0 = reinterpret_cast<const uint32_t*>(wq + (size_t)r0 * p.Cin + k0);
                    const uint32_t* w1 = reinterpret_cast<const uint32_t*>(wq + (size_t)r1 * p.Cin + k0);
                    uint32_t af[4];
                    af[0] = (v0 && r0 < p.Cout) ? w0[0] : 0u;
                    af[1] = (v0 && r1 < p.Cout) ? w1[0] : 0u;
                    af[2] = (v1 && r0 < p.Cout) ? w0[4] : 0u;
                    af[3] = (v1 && r1 < p.Cout) ? w1[4] : 0u;
#pragma unroll
                    for (int j = 0; j < NTP; ++j) mma_bf16_16816(acc[m][j], af, bf[j]);
                }
            }
            // D fragment: rows g / g + 8 of the Cout tile, columns 2t, 2t + 1 of the pixel tile (neighbours in one patch row when pw is even)
            __nv_bfloat16* ypatch = p.y + ((size_t)b * p.Cout * p.H + (size_t)pi * p.ph) * p.W + (size_t)jg * PGW + q * p.pw;
#pragma unroll
            for (int j = 0; j < NTP; ++j) {
                const int pp = j * 8 + 2 * t;
                if (pp >= npix) continue;
                const int pr = pp / p.pw, pc = pp % p.pw, pr1 = (pp + 1) / p.pw, pc1 = (pp + 1) % p.pw;
                const bool pair = (p.pw & 1) == 0;              // pixels pp, pp + 1 are neighbours in one patch row
#pragma unroll
                for (int m = 0; m < MT; ++m)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int oc = m * 16 + h * 8 + g;
                        if (oc >= p.Cout) continue;
                        const float a0 = act_apply(fmaf(acc[m][j][2 * h], sc[m][h], sh[m][h]), p.act);
                        const float a1 = act_apply(fmaf(acc[m][j][2 * h + 1], sc[m][h], sh[m][h]), p.act);
                        __nv_bfloat16* dst = ypatch + (size_t)oc * ostride + (size_t)pr * p.W + pc;
                        if (pair) *reinterpret_cast<__nv_bfloat162*>(dst) = __floats2bfloat162_rn(a0, a1);
                        else {
                            *dst = __float2bfloat16_rn(a0);
                            if (pp + 1 < npix) ypatch[(size_t)oc * ostride + (size_t)pr1 * p.W + pc1] = __float2bfloat16_rn(a1);
                        }
                    }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty_w + s * MMA_MAX_SLOTS + q);      // the weight row of this slot may be overwritten
        }
        __syncwarp();
        if (lane == 0 && warp < users) mbar_arrive(empty_x + s);              // this warp is done with the unit's x tile
    }
}

template <int MT, int NTP>
static int launch_mma(const CUtensorMap& xmap, const RingParams& p, size_t smem, int grid, cudaStream_t st) {
    auto k = conv1x1_mma_kernel<MT, NTP>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(HSB_ERR_CUDA, std::string("conv1x1_mma attr: ") + cudaGetErrorString(e));
    k<<<grid, MMA_THREADS, smem, st>>>(xmap, p);
    note_kernel("conv1x1_mma_kernel");
    return check_launch("conv1x1_mma launch");
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
    if (ph > 256 || ((uintptr_t)x & 15) || ((uintptr_t)w & 15) || ((uintptr_t)y & 3) || (W * 2) % 16 || (hp * 2) % 16 || (w_row_stride * 2) % 16) return HSB_OK;
    EncodeTiledFnR encode = ring_encode_fn();
    if (!encode) return HSB_OK;
    const int wrow_e = ((Cin * Cout + 7) / 8 * 8) + ((((Cin * Cout + 7) / 8) % 2 == 0) ? 8 : 0);
    const cuuint64_t dim[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)Cin, (cuuint64_t)B};
    const cuuint64_t str[3] = {(cuuint64_t)W * 2, (cuuint64_t)H * W * 2, (cuuint64_t)Cin * H * W * 2};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    static const bool verbose = [] { const char* v = getenv("HSB_VERBOSE"); return v && v[0] == '1'; }();

    // ---- tensor-core variant: patches of up to 16 pixels, up to 64 output channels.  Opt-in (HSB_CONV_MMA=1, read at every
    // call so that tests can switch it): on B200 it measured 16.0 / 12.4 / 10.8 us at the three HyperSeg-M levels against
    // 14.3 (CUDA-core ring) / 12.6 / 10.3 us (one-shot kernel) -- these launches are bound by launch + first-load latency, not
    // by arithmetic or by the ring's granularity, so the simpler kernels stay the default. ----
    const char* mma_env = getenv("HSB_CONV_MMA");
    if (mma_env && mma_env[0] == '1' && !all && ph * pw <= 16 && Cout <= 64) {
        // unit = PG patches, one per warp where possible: the smallest valid PG >= 8, else the largest valid one
        int PG = 0;
        for (int cand = 1; cand <= fw && cand <= MMA_MAX_SLOTS; ++cand) {
            if (fw % cand || (cand * pw * 2) % 16 || cand * pw > 256) continue;
            const size_t stage = (size_t)cand * wrow_e * 2 + (size_t)Cin * ph * cand * pw * 2 + 256;
            if (2 * stage > 200 * 1024) break;
            PG = cand;
            if (cand >= 8) break;
        }
        if (PG) {
            const int PGW = PG * pw, plane = ph * PGW;
            RingParams p;
            p.w = (const __nv_bfloat16*)w; p.y = (__nv_bfloat16*)y; p.post_scale = post_scale; p.post_shift = post_shift; p.act = act;
            p.B = B; p.Cin = Cin; p.Cout = Cout; p.H = H; p.W = W; p.fh = fh; p.fw = fw; p.ph = ph; p.pw = pw;
            p.PG = PG; p.upr = fw / PG; p.wsplit = 1; p.PGw = PG; p.units = B * fh * p.upr; p.hp = hp; p.w_row_stride = w_row_stride;
            p.wrow = wrow_e;
            p.x_off = (int)(((size_t)PG * p.wrow * 2 + 127) / 128 * 128);
            p.stage_bytes = (int)(((size_t)p.x_off + (size_t)Cin * plane * 2 + 127) / 128 * 128);
            p.stages = std::min(MMA_MAX_STAGES, (int)((220 * 1024) / p.stage_bytes));
            p.lanes_over_pixels = 0;
            const size_t smem = MMA_BAR_BYTES + (size_t)p.stages * p.stage_bytes;
            CUtensorMap xmap;
            const cuuint32_t box[4] = {(cuuint32_t)PGW, (cuuint32_t)ph, (cuuint32_t)Cin, 1};
            if (p.stages >= 2 &&
                encode(&xmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dim, str, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS) {
                const int grid = std::min(p.units, std::max(1, device_sm_count()));
                const int MT = (Cout + 15) / 16, NTP = (ph * pw + 7) / 8;
                if (verbose)
                    fprintf(stderr, "[hsb] conv1x1_mma: %d -> %d, %dx%d px/patch: %d patches per unit, %d units, %d Cout tiles x %d pixel tiles, %d stages of %d B\n",
                            Cin, Cout, ph, pw, PG, p.units, MT, NTP, p.stages, p.stage_bytes);
                *handled = true;
#define HSB_MMA_CASE(M, N) if (MT == M && NTP == N) return launch_mma<M, N>(xmap, p, smem, grid, st);
                HSB_MMA_CASE(1, 1) HSB_MMA_CASE(2, 1) HSB_MMA_CASE(3, 1) HSB_MMA_CASE(4, 1)
                HSB_MMA_CASE(1, 2) HSB_MMA_CASE(2, 2) HSB_MMA_CASE(3, 2) HSB_MMA_CASE(4, 2)
#undef HSB_MMA_CASE
                *handled = false;
            }
        }
    }
    if (ph * pw > 1 && !all) return HSB_OK;
    // patches per unit: the x box must have 16-byte rows (PG * pw * 2 bytes) and PG must divide fw; the smallest such PG that
    // gives every thread an item (units stay small: more units than SMs, several stages in flight), else the largest one
    const int PB = (pw % 2 == 0) ? 2 : 1;
    static const int env_threads = [] { const char* v = getenv("HSB_RING_THREADS"); return v ? atoi(v) : 0; }();
    static const int env_items = [] { const char* v = getenv("HSB_RING_ITEMS"); return v ? atoi(v) : 0; }();
    const int threads = env_threads >= 64 && env_threads <= RING_MAX_THREADS && env_threads % 32 == 0 ? env_threads : 256;
    const int min_items = env_items > 0 ? env_items : threads;
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
    CUtensorMap xmap;
    const cuuint32_t box[4] = {(cuuint32_t)PGW, (cuuint32_t)ph, (cuuint32_t)Cin, 1};
    if (encode(&xmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dim, str, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return HSB_OK;
    const int grid = std::min(p.units, std::max(1, device_sm_count()));
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
