// Patch-wise 1x1 convolution with fused BN(eval)+activation epilogue.
//
// Replaces HyperPatchNoPadding.forward (reference hyperseg/models/hyperseg_v1_0.py:486-498) together
// with the BatchNorm2d + ReLU appended by make_hyper_patch_conv2d_block (:753-756).
//
// The op is bound by streaming the per-patch weight rows (at the coarse decoder levels one patch is
// 1..16 pixels but 700..5000 weights), so the kernel is organised around that stream:
//   * one CTA owns PG consecutive patches of one patch-row (same image, same i);
//   * the PG weight rows are pulled into shared memory by the TMA engine (cp.async.bulk, one
//     request per row, byte-counted on an mbarrier) while all threads stage the input pixels;
//   * each thread then produces an OB x PB register tile (OB output channels, PB adjacent pixels),
//     reading weights as packed pairs along Cin;
//   * the epilogue applies scale/shift/activation and writes NCHW rows (adjacent lanes = adjacent
//     pixels, so stores coalesce along W).
// Several CTAs are resident per SM (<= ~48 KB shared memory each), so one CTA's weight stream
// overlaps another's arithmetic.
#include "bulk_copy.cuh"
#include "common.cuh"

namespace hsb {

struct Conv1x1Params {
    const void* x; const void* w; void* y;
    const float* post_scale; const float* post_shift; int act;
    int B, Cin, Cout, H, W, fh, fw, ph, pw, groups, cig, cog;
    WStrides ws;
    int hp;            // weights per patch
    int PG;            // patches per CTA
    int jgroups;       // ceil(fw / PG)
    int wrow_smem;     // elements between staged weight rows (multiple of 8)
    int bulk_ok;       // weight rows satisfy the cp.async.bulk alignment rules
};

template <typename T> struct Pair;
template <> struct Pair<float> {
    __device__ static float2 load(const float* p) { return *reinterpret_cast<const float2*>(p); }
};
template <> struct Pair<__nv_bfloat16> {
    __device__ static float2 load(const __nv_bfloat16* p) {
        return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
    }
};

__device__ __forceinline__ void store_pair(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
__device__ __forceinline__ void store_pair(__nv_bfloat16* p, float a, float b) {
    *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}

// OB: output channels per thread, PB: adjacent pixels per thread (1 or 2), PAIRC: Cin/G is even so
// weights can be read as (c, c+1) pairs.
template <typename T, int OB, int PB, bool PAIRC>
__global__ void __launch_bounds__(128) patch_conv1x1_kernel(const Conv1x1Params p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    T* wsm = reinterpret_cast<T*>(smem_raw + 16);
    const int tid = threadIdx.x;

    const int jg = blockIdx.x % p.jgroups;
    const int bi = blockIdx.x / p.jgroups;         // b*fh + i
    const int b = bi / p.fh, pi = bi % p.fh;
    const int j0 = jg * p.PG;
    const int npat = min(p.PG, p.fw - j0);         // patches actually present in this CTA
    const int PGW = p.PG * p.pw;                   // staged row pitch (pixels)
    const int ncol = npat * p.pw;                  // valid pixels per staged row
    T* xsm = wsm + (size_t)p.PG * p.wrow_smem;     // [Cin][ph][PGW]

    const T* x = reinterpret_cast<const T*>(p.x);
    const T* w = reinterpret_cast<const T*>(p.w);
    T* y = reinterpret_cast<T*>(p.y);
    const T* wbase = w + (size_t)b * p.ws.b + (size_t)(pi * p.fw + j0) * p.ws.p;

    // ---- stage weights ------------------------------------------------------------------
    if (p.bulk_ok) {
        if ((tid >> 5) == 0 && elect_one()) {
            mbar_init(bar, 1);
            mbar_fence_init();
            const uint32_t row_bytes = (uint32_t)p.hp * sizeof(T);
            mbar_arrive_expect_tx(bar, row_bytes * npat);
            for (int q = 0; q < npat; ++q)
                bulk_g2s(wsm + (size_t)q * p.wrow_smem, wbase + (size_t)q * p.ws.p, row_bytes, bar);
        }
    } else {
        for (int q = 0; q < npat; ++q) {
            const T* src = wbase + (size_t)q * p.ws.p;
            T* dst = wsm + (size_t)q * p.wrow_smem;
            for (int k = tid; k < p.hp; k += blockDim.x) dst[k] = src[(size_t)k * p.ws.k];
        }
    }

    // ---- stage input pixels: rows of ncol contiguous elements per (channel, patch row) ----
    {
        const T* xrow0 = x + (((size_t)b * p.Cin) * p.H + (size_t)pi * p.ph) * p.W + (size_t)j0 * p.pw;
        const int rows = p.Cin * p.ph;
        const bool vec = (sizeof(T) * ncol) % 16 == 0 && (sizeof(T) * PGW) % 16 == 0 &&
                         (sizeof(T) * p.W) % 16 == 0 && aligned16(xrow0) && aligned16(xsm);
        if (vec) {
            const int cpr = (int)(sizeof(T) * ncol / 16);          // 16-byte chunks per row
            for (int idx = tid; idx < rows * cpr; idx += blockDim.x) {
                int row = idx / cpr, ch = idx % cpr;
                int c = row / p.ph, r = row % p.ph;
                const uint4* src = reinterpret_cast<const uint4*>(xrow0 + ((size_t)c * p.H + r) * p.W) + ch;
                uint4* dst = reinterpret_cast<uint4*>(xsm + (size_t)row * PGW) + ch;
                *dst = __ldg(src);
            }
        } else {
            for (int idx = tid; idx < rows * ncol; idx += blockDim.x) {
                int row = idx / ncol, q = idx % ncol;
                int c = row / p.ph, r = row % p.ph;
                xsm[(size_t)row * PGW + q] = xrow0[((size_t)c * p.H + r) * p.W + q];
            }
        }
    }
    __syncthreads();                       // xsm (and non-bulk wsm) visible; barrier init visible
    if (p.bulk_ok) mbar_wait(bar, 0);      // weight rows have landed

    // ---- compute ------------------------------------------------------------------------
    const int ogs = p.Cout / OB;                       // launcher guarantees cog % OB == 0
    const int qcols = PGW / PB;                        // pixel groups per staged row (PGW % PB == 0)
    const int items = ogs * p.ph * qcols;
    const int plane = p.ph * PGW;
    for (int it = tid; it < items; it += blockDim.x) {
        const int qc = it % qcols;
        const int r = (it / qcols) % p.ph;
        const int og = it / (qcols * p.ph);
        const int col = qc * PB;
        if (col >= ncol) continue;
        const bool second = (PB == 2) && (col + 1 < ncol);   // ragged last CTA with odd pw
        const int pp = col / p.pw;
        const int o0 = og * OB;
        const int g = o0 / p.cog;
        const T* wrow = wsm + (size_t)pp * p.wrow_smem + (size_t)o0 * p.cig;
        const T* xc = xsm + (size_t)(g * p.cig) * plane + r * PGW + col;
        float acc[OB][PB];
#pragma unroll
        for (int k = 0; k < OB; ++k)
#pragma unroll
            for (int q = 0; q < PB; ++q) acc[k][q] = 0.f;

        int c = 0;
        if (PAIRC) {
            for (; c + 1 < p.cig; c += 2) {
                float xa[PB], xb[PB];
                if constexpr (PB == 2) {
                    float2 t0 = Pair<T>::load(xc + (size_t)c * plane);
                    float2 t1 = Pair<T>::load(xc + (size_t)(c + 1) * plane);
                    xa[0] = t0.x; xa[PB - 1] = t0.y; xb[0] = t1.x; xb[PB - 1] = t1.y;
                } else {
                    xa[0] = ld_f(xc + (size_t)c * plane);
                    xb[0] = ld_f(xc + (size_t)(c + 1) * plane);
                }
#pragma unroll
                for (int k = 0; k < OB; ++k) {
                    float2 wv = Pair<T>::load(wrow + k * p.cig + c);
#pragma unroll
                    for (int q = 0; q < PB; ++q) {
                        acc[k][q] = fmaf(wv.x, xa[q], acc[k][q]);
                        acc[k][q] = fmaf(wv.y, xb[q], acc[k][q]);
                    }
                }
            }
        }
        for (; c < p.cig; ++c) {
            float xa[PB];
#pragma unroll
            for (int q = 0; q < PB; ++q) xa[q] = ld_f(xc + (size_t)c * plane + q);
#pragma unroll
            for (int k = 0; k < OB; ++k) {
                float wv = ld_f(wrow + k * p.cig + c);
#pragma unroll
                for (int q = 0; q < PB; ++q) acc[k][q] = fmaf(wv, xa[q], acc[k][q]);
            }
        }

        // epilogue
        T* yrow = y + (((size_t)b * p.Cout + o0) * p.H + (size_t)pi * p.ph + r) * p.W + (size_t)j0 * p.pw + col;
        const size_t ostride = (size_t)p.H * p.W;
#pragma unroll
        for (int k = 0; k < OB; ++k) {
            float v[PB];
#pragma unroll
            for (int q = 0; q < PB; ++q) {
                float t = acc[k][q];
                if (p.post_scale) t = fmaf(t, p.post_scale[o0 + k], p.post_shift[o0 + k]);
                v[q] = act_apply(t, p.act);
            }
            T* dst = yrow + k * ostride;
            if constexpr (PB == 2) {
                if (second && (p.W & 1) == 0) store_pair(dst, v[0], v[PB - 1]);
                else { st_f(dst, v[0]); if (second) st_f(dst + 1, v[PB - 1]); }
            } else {
                st_f(dst, v[0]);
            }
        }
    }
}

template <typename T, int OB, int PB>
static int launch_1x1(const Conv1x1Params& p, size_t smem, cudaStream_t st) {
    dim3 grid(p.B * p.fh * p.jgroups);
    cudaError_t e;
    if (p.cig % 2 == 0) {
        auto k = patch_conv1x1_kernel<T, OB, PB, true>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(HSB_ERR_CUDA, std::string("conv1x1 attr: ") + cudaGetErrorString(e));
        cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        k<<<grid, 128, smem, st>>>(p);
    } else {
        auto k = patch_conv1x1_kernel<T, OB, PB, false>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(HSB_ERR_CUDA, std::string("conv1x1 attr: ") + cudaGetErrorString(e));
        cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        k<<<grid, 128, smem, st>>>(p);
    }
    note_kernel("patch_conv1x1_kernel");
    return check_launch("patch_conv1x1 launch");
}

template <typename T>
static int dispatch_1x1(Conv1x1Params& p, cudaStream_t st) {
    const size_t es = sizeof(T);
    const int pix = p.ph * p.pw;
    // patches per CTA: at most ~44 KB of shared memory (several CTAs per SM overlap load and compute), a divisor
    // of fw when one is close, and small enough that the grid has >= 4 CTAs per SM
    const size_t per_patch = (size_t)p.wrow_smem * es + (size_t)p.Cin * pix * es;
    int PG = (int)std::max<size_t>(1, (44 * 1024) / per_patch);
    PG = std::min(PG, p.fw);
    PG = std::min(PG, 32);
    const int64_t rows = (int64_t)p.B * p.fh;
    const int64_t want = 4LL * std::max(1, device_sm_count());
    while (PG > 1 && rows * ceil_div(p.fw, PG) < want) --PG;
    for (int cand = PG; cand >= std::max(1, PG / 2); --cand)
        if (p.fw % cand == 0) { PG = cand; break; }
    int PB = ((PG * p.pw) % 2 == 0 && pix >= 4) ? 2 : 1;
    p.PG = PG;
    p.jgroups = ceil_div(p.fw, PG);
    const size_t smem = 16 + ((size_t)PG * p.wrow_smem + (size_t)p.Cin * pix * PG) * es;
    HSB_REQUIRE(smem <= 220 * 1024, HSB_ERR_UNSUPPORTED,
                "patch_conv1x1: one patch needs " + std::to_string(smem) + " B of shared memory");
    const int OB = (p.cog % 4 == 0) ? 4 : (p.cog % 2 == 0 ? 2 : 1);
    if (PB == 2) {
        if (OB == 4) return launch_1x1<T, 4, 2>(p, smem, st);
        if (OB == 2) return launch_1x1<T, 2, 2>(p, smem, st);
        return launch_1x1<T, 1, 2>(p, smem, st);
    }
    if (OB == 4) return launch_1x1<T, 4, 1>(p, smem, st);
    if (OB == 2) return launch_1x1<T, 2, 1>(p, smem, st);
    return launch_1x1<T, 1, 1>(p, smem, st);
}

int conv1x1_ring_try(const void* x, const void* w, void* y, const float* post_scale, const float* post_shift, int act, int B, int Cin,
                     int Cout, int H, int W, int fh, int fw, int groups, int64_t w_row_stride, cudaStream_t st, bool* handled);

}  // namespace hsb

using namespace hsb;

extern "C" int hsb_patch_conv1x1_fwd(const void* x, const void* w, void* y,
                                     const float* post_scale, const float* post_shift, int act,
                                     int B, int Cin, int Cout, int H, int W, int fh, int fw, int groups,
                                     int dtype, int w_layout, int64_t w_row_stride, void* stream) {
    HSB_REQUIRE(x && w && y, HSB_ERR_INVALID_ARG, "patch_conv1x1: null pointer");
    HSB_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && H > 0 && W > 0 && fh > 0 && fw > 0 && groups > 0,
                HSB_ERR_INVALID_ARG, "patch_conv1x1: non-positive dimension");
    HSB_REQUIRE(H % fh == 0 && W % fw == 0, HSB_ERR_INVALID_ARG,
                "patch_conv1x1: feature map is not divisible into fh x fw patches");
    HSB_REQUIRE(Cin % groups == 0 && Cout % groups == 0, HSB_ERR_INVALID_ARG,
                "patch_conv1x1: channels not divisible by groups");
    HSB_REQUIRE((post_scale == nullptr) == (post_shift == nullptr), HSB_ERR_INVALID_ARG,
                "patch_conv1x1: post_scale and post_shift must be given together");
    HSB_REQUIRE(dtype == HSB_F32 || dtype == HSB_BF16, HSB_ERR_INVALID_ARG, "patch_conv1x1: bad dtype");
    HSB_REQUIRE(act >= HSB_ACT_NONE && act <= HSB_ACT_RELU6, HSB_ERR_INVALID_ARG, "patch_conv1x1: bad act");
    Conv1x1Params p;
    p.x = x; p.w = w; p.y = y; p.post_scale = post_scale; p.post_shift = post_shift; p.act = act;
    p.B = B; p.Cin = Cin; p.Cout = Cout; p.H = H; p.W = W; p.fh = fh; p.fw = fw;
    p.ph = H / fh; p.pw = W / fw; p.groups = groups; p.cig = Cin / groups; p.cog = Cout / groups;
    p.hp = Cout * p.cig;
    if (w_layout == HSB_W_PATCH_MAJOR)
        HSB_REQUIRE(w_row_stride >= p.hp, HSB_ERR_INVALID_ARG, "patch_conv1x1: w_row_stride < hyper params");
    p.ws = make_wstrides(w_layout, p.hp, (int64_t)fh * fw, w_row_stride);
    const size_t es = dtype == HSB_F32 ? 4 : 2;
    // staged weight rows start 16-byte aligned; an odd number of 16-byte units per row makes consecutive patches
    // land in different bank groups (lanes of one warp read the same offset of different patches' rows)
    p.wrow_smem = (p.hp + 7) / 8 * 8;
    if (((size_t)p.wrow_smem * es / 16) % 2 == 0) p.wrow_smem += (int)(16 / es);
    p.bulk_ok = (w_layout == HSB_W_PATCH_MAJOR) && ((uintptr_t)w % 16 == 0) &&
                ((size_t)p.hp * es) % 16 == 0 && ((size_t)w_row_stride * es) % 16 == 0;
    HSB_REQUIRE((int64_t)B * fh * fw < (1ll << 31), HSB_ERR_UNSUPPORTED, "patch_conv1x1: too many patches");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == HSB_BF16 && w_layout == HSB_W_PATCH_MAJOR) {       // the decoder's case: persistent ring kernel when it fits
        bool handled = false;
        const int rc = conv1x1_ring_try(x, w, y, post_scale, post_shift, act, B, Cin, Cout, H, W, fh, fw, groups, w_row_stride, st, &handled);
        if (handled) return rc;
    }
    if (dtype == HSB_F32) return dispatch_1x1<float>(p, st);
    return dispatch_1x1<__nv_bfloat16>(p, st);
}
