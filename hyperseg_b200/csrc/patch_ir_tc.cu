// Fused patch-wise inverted-residual MetaBlock -- bf16 tensor-core path (tcgen05 / TMEM / TMA), 16x16 and 8x8 patches, for
// per-patch weights handed over in the REFERENCE order (hsb_patch_ir_fwd): this kernel re-stages them (and the x tile) into
// UMMA operands itself.  When the weights come from our own head, the restage-free kernel in patch_ir2.cu is used instead.
//
// Same arithmetic as patch_ir.cu (reference hyperseg/models/hyperseg_v1_0.py:328-376), organised for sm_100a:
//
//   persistent CTAs, TWO (16x16) or THREE (8x8) resident per SM so that one CTA's barrier / tensor-core / TMA
//   latencies are covered by another CTA's arithmetic; each CTA walks its patches:
//   P0  wait for the patch's weight row (cp.async.bulk, own mbarrier: it was requested a whole patch ago);
//   P1  re-stage W1/W3 -> UMMA operands B1/B2 with the BatchNorm scale folded in and the shift as an extra K column,
//       W2 -> packed bf16x2 taps; only then wait for the x tile -- the (ph+2)x(pw+2) halo tile arrives through a 4-D
//       tensor map (box starting 8 pixels left of the patch: TMA wants a 16-byte aligned innermost start; zero fill
//       outside the image), so its flight time hides behind the weight re-stage -- patch the reflect halo of
//       image-border patches in place and re-stage x -> operand A1 (K-major, lanes over pixels: conflict-free, with a
//       constant-one channel that carries the BatchNorm shift);
//   P2  GEMM1 on the tensor core: H[(ph+2)(pw+2) px x hid] = A1 . B1^T, M=128 tiles accumulated in TMEM, issued by an
//       elected lane of warp 0 while an elected lane of warp 1 requests the next patch's weight row;
//   P3  epilogue 1, tile by tile as the MMAs retire: TMEM -> registers -> ReLU6 -> bf16 -> shared "hidden" tile
//       [pixel][channel];
//   P4  depthwise 3x3 + BN2 + ReLU6 on CUDA cores in packed bf16x2 (lane = channel pair, warp = tile column,
//       3x3 register window sliding down the column), written straight into GEMM2's A operand
//       (128B-swizzled K-major);
//   P5  GEMM2: O[ph*pw px x Cout] = A2 . B2^T;
//   P6  epilogue 2: TMEM -> registers -> bf16 -> NCHW global stores (the next x tile has been in flight since P5).
//
// Shared memory is time-shared inside a patch so that two patches-in-flight fit an SM (~110 KB per CTA at the level-4
// shape): region Y holds A1+B1 during P1-P2 and A2 during P4-P5; region X holds the raw x tile during P0-P1 and the
// hidden tile during P3-P4, so the next x tile can be requested as soon as the depthwise phase has consumed the hidden tile.  Stale bytes left behind by the other tenant are finite bf16 values that only ever meet
// zero weights (padding K columns) or land in accumulator rows nobody reads.  TMEM is time-shared the same way
// (GEMM2's accumulators reuse GEMM1's columns).
// The BatchNorm shifts ride inside the GEMMs (constant-one K column), so the epilogues are a clamp and a convert.
// Nothing but x, the weight row and y touches global memory.
#include <cuda.h>

#include <cstdio>
#include <mutex>

#include "common.cuh"
#include "tcgen05.cuh"

namespace hsb {

constexpr int r16(int v) { return (v + 15) / 16 * 16; }
constexpr int r8(int v) { return (v + 7) / 8 * 8; }
constexpr int r128(int v) { return (v + 127) / 128 * 128; }
constexpr int pow2_at_least(int v) { int p = 32; while (p < v) p *= 2; return p; }
constexpr int imax(int a, int b) { return a > b ? a : b; }
constexpr int imin(int a, int b) { return a < b ? a : b; }

template <int CIN_, int HID_, int COUT_, int PS_>
struct IRTC {
    static constexpr int CIN = CIN_, HID = HID_, COUT = COUT_;
    static constexpr int PH = PS_, PW = PS_, TH = PS_ + 2, TW = PS_ + 2;
    // TMA needs a 16-byte aligned start in the innermost dimension: the box starts 8 pixels left of the patch,
    // the halo tile's column 0 is box column XOFF.
    static constexpr int XOFF = 7, TWB = r8(XOFF + TW);
    static constexpr int T = TH * TW, O = PH * PW;
    static constexpr int K1 = r16(CIN + 1), N1 = r16(HID), K2 = r16(HID + 1), N2 = r16(COUT);
    static constexpr int M1T = (T + 127) / 128, M2T = (O + 127) / 128;
    static constexpr int MC1 = M1T * 16, MC2 = M2T * 16;   // 8-pixel chunks of the A operands
    static constexpr int HP = CIN * HID + 9 * HID + HID * COUT;
    static constexpr int R1 = CIN * HID, R2 = R1 + 9 * HID;
    static constexpr int NPAIR = HID / 2;
    static constexpr int MAINP = NPAIR < 32 ? NPAIR : 32;
    static constexpr int TAILP = NPAIR - MAINP;            // channel pairs living in the K >= 64 tail
    static constexpr int HPITCH = r8(HID);                 // hidden tile pitch (elements), 16-byte multiple
    static constexpr int HPW = HPITCH / 2;                 // ... in 32-bit words
    static constexpr int KT2 = K2 > 64 ? K2 - 64 : 0;      // K extent of A2's non-swizzled tail
    static constexpr int DWW = 8;                          // warps that walk tile columns in the depthwise phase
    static constexpr int WARPS = DWW + (TAILP > 0 ? 1 : 0);
    static constexpr int THREADS = 32 * WARPS;
    static constexpr int TMEM_COLS = pow2_at_least(imax(M1T * N1, M2T * N2));   // GEMM2 reuses GEMM1's columns
    // UMMA operand strides (bytes)
    static constexpr int A1_SBO = 128, A1_LBO = MC1 * 128;
    static constexpr int B1_SBO = 128, B1_LBO = (N1 / 8) * 128;
    static constexpr int B2_SBO = 128, B2_LBO = (N2 / 8) * 128;
    static constexpr int A2T_SBO = 128, A2T_LBO = MC2 * 128;
    // shared memory map (bytes from a 1024-aligned base)
    static constexpr int SZ_A2 = M2T * 128 * 128, SZ_A2T = (KT2 / 8) * MC2 * 128;
    static constexpr int SZ_RAWX = CIN * TH * TWB * 2;
    static constexpr int SZ_A1 = (K1 / 8) * MC1 * 128, SZ_B1 = (K1 / 8) * (N1 / 8) * 128;
    static constexpr int SZ_HID = r128(T * HPITCH * 2);
    static constexpr int OFF_Y = 0;                                             // region Y: A1 + B1 (P1-P2) | A2 (+ tail) (P4-P5)
    static constexpr int SZ_Y = r128(imax(SZ_A1 + SZ_B1, SZ_A2 + SZ_A2T));
    static constexpr int OFF_A2 = OFF_Y, OFF_A2T = OFF_Y + SZ_A2, OFF_A1 = OFF_Y, OFF_B1 = OFF_Y + SZ_A1;
    static constexpr int OFF_X = OFF_Y + SZ_Y;                                  // region X: raw x tile (P0-P1) | hidden tile (P3-P4)
    static constexpr int SZ_X = r128(imax(SZ_RAWX, SZ_HID));
    static constexpr int OFF_RAWX = OFF_X, OFF_HID = OFF_X;
    static constexpr int OFF_B2 = OFF_X + SZ_X;
    static constexpr int SZ_B2 = (K2 / 8) * (N2 / 8) * 128;
    static constexpr int OFF_RAWW = OFF_B2 + SZ_B2;
    static constexpr int SZ_RAWW = r16(HP * 2) + 16;
    static constexpr int OFF_W2P = r16(OFF_RAWW + SZ_RAWW);
    static constexpr int SZ_W2P = 10 * HPW * 4;
    static constexpr int OFF_BN = OFF_W2P + SZ_W2P;
    static constexpr int SZ_BN = (4 * HID + 2 * COUT) * 4;
    static constexpr int OFF_BAR = r16(OFF_BN + SZ_BN);
    static constexpr int OFF_COORD = OFF_BAR + 64;            // int[2][4]: (b, pi, pj) of the patch in flight
    static constexpr int USED_BYTES = OFF_COORD + 32;
    static constexpr int SMEM_BYTES = USED_BYTES + 1024;      // + slack for the 1024-byte alignment
    // CTAs per SM: bounded by shared memory, TMEM columns and the register file
    static constexpr int CTAS = imax(1, imin(imin(227 * 1024 / SMEM_BYTES, 512 / TMEM_COLS), PS_ == 16 ? 2 : 3));
    static_assert(HID % 4 == 0 && HID <= 68, "hidden width must be a multiple of 4 and at most 68");
    static_assert(TAILP <= 2, "at most two channel pairs in the K tail");
    static_assert(PW % DWW == 0, "tile columns must divide among the depthwise warps");
    static_assert(TMEM_COLS <= 512 && 1 + M1T + M2T <= 7, "TMEM / barrier budget");
    static_assert(PS_ == 8 || PS_ == 16, "patch size");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

struct IRTCParams {
    const __nv_bfloat16* w;
    __nv_bfloat16* y;
    const float* bn[6];
    int B, H, W, fh, fw;
    int64_t w_row_stride;      // elements between patch rows
    int w_bulk;                // rows are 16-byte aligned -> cp.async.bulk
    int total;                 // B * fh * fw
#ifdef HSB_IR_PROF
    long long* prof;           // [grid][10] per-phase cycle sums (profiling build only, scripts/ir_phase_prof.sh)
#endif
};

__device__ __forceinline__ uint32_t pack_relu6(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    const __nv_bfloat162 six = __floats2bfloat162_rn(6.f, 6.f);
    __nv_bfloat162 v = __hmin2(*reinterpret_cast<__nv_bfloat162*>(&r), six);
    return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
}

#ifdef HSB_IR_PROF
#define HSB_STAMP(k) do { if (tid == 0) { long long now_ = clock64(); prof_acc[k] += now_ - prof_t; prof_t = now_; } } while (0)
#else
#define HSB_STAMP(k) do { } while (0)
#endif

template <class C>
__global__ void __launch_bounds__(C::THREADS, C::CTAS)
patch_ir_tc_kernel(const __grid_constant__ CUtensorMap xmap, const IRTCParams p) {
    extern __shared__ unsigned char smem_dyn[];
    // align by pointer arithmetic on the __shared__ array so the compiler keeps the address space (LDS/STS, not generic)
    unsigned char* sm = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    __nv_bfloat16* rawX = reinterpret_cast<__nv_bfloat16*>(sm + C::OFF_RAWX);
    __nv_bfloat16* rawW = reinterpret_cast<__nv_bfloat16*>(sm + C::OFF_RAWW);
    uint32_t* w2p = reinterpret_cast<uint32_t*>(sm + C::OFF_W2P);            // [10][HPW] bf16x2
    float* bnsm = reinterpret_cast<float*>(sm + C::OFF_BN);
    float* s1 = bnsm, *b1 = s1 + C::HID, *s2 = b1 + C::HID, *b2 = s2 + C::HID, *s3 = b2 + C::HID, *b3 = s3 + C::COUT;
    uint64_t* bar_tma = reinterpret_cast<uint64_t*>(sm + C::OFF_BAR);
    uint64_t* bar_mma1 = bar_tma + 1;                     // [M1T] one per GEMM1 tile
    uint64_t* bar_mma2 = bar_mma1 + C::M1T;               // [M2T]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma2 + C::M2T);
    uint64_t* bar_w = bar_tma + 7;                        // weight row of the patch (bar_tma then counts the x tile only)
    volatile int* coord = reinterpret_cast<volatile int*>(sm + C::OFF_COORD);

    const uint32_t a1_addr = smem_u32(sm + C::OFF_A1), b1_addr = smem_u32(sm + C::OFF_B1);
    const uint32_t a2_addr = smem_u32(sm + C::OFF_A2), a2t_addr = smem_u32(sm + C::OFF_A2T);
    const uint32_t b2_addr = smem_u32(sm + C::OFF_B2);

    // ---------------- one-time setup ----------------
    // every byte that the tensor core may read as padding must be a finite value: clear the whole arena once
    for (int i = tid; i < C::OFF_BAR / 16; i += C::THREADS) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    for (int i = tid; i < C::HID; i += C::THREADS) {
        s1[i] = p.bn[0][i]; b1[i] = p.bn[1][i]; s2[i] = p.bn[2][i]; b2[i] = p.bn[3][i];
    }
    for (int i = tid; i < C::COUT; i += C::THREADS) { s3[i] = p.bn[4][i]; b3[i] = p.bn[5][i]; }
    if (tid == 0) {
        mbar_init(bar_tma, 1);
        mbar_init(bar_w, 1);
        for (int t = 0; t < C::M1T; ++t) mbar_init(bar_mma1 + t, 1);
        for (int t = 0; t < C::M2T; ++t) mbar_init(bar_mma2 + t, 1);
        mbar_fence_init();
        tma_prefetch_desc(&xmap);
    }
    if (warp == 0) tmem_alloc(tmem_slot, C::TMEM_COLS);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    const int P = p.fh * p.fw;
    constexpr uint32_t X_BYTES = C::SZ_RAWX;
    constexpr uint32_t W_BYTES = r16(C::HP * 2);
    // A patch is announced (coordinates) when its weight row is requested; the x tile is requested later, once its
    // landing buffer (region X) is free.  The weight row completes on bar_w, the x tile on bar_tma, so the weight
    // re-stage of P1 runs while the x tile is still in flight.
    auto announce_and_load_w = [&](int patch, uint32_t slot) {         // one thread
        const int b = patch / P, pp = patch % P, pi = pp / p.fw, pj = pp % p.fw;
        coord[slot * 4 + 0] = b; coord[slot * 4 + 1] = pi; coord[slot * 4 + 2] = pj;   // published by the arrive below
        if (p.w_bulk) {
            mbar_arrive_expect_tx(bar_w, W_BYTES);
            bulk_g2s(rawW, p.w + (size_t)patch * p.w_row_stride, W_BYTES, bar_w);
        }
    };
    auto load_x = [&](uint32_t slot) {                                 // one thread, after the announcement is visible to it
        mbar_arrive_expect_tx(bar_tma, X_BYTES);
        tma_load_4d(rawX, &xmap, coord[slot * 4 + 2] * C::PW - 8, coord[slot * 4 + 1] * C::PH - 1, 0, coord[slot * 4 + 0],
                    bar_tma);
    };
    // in the loop an elected lane of warp 1 issues the loads while an elected lane of warp 0 issues the MMAs
    if (tid == 0 && (int)blockIdx.x < p.total) {
        announce_and_load_w(blockIdx.x, 0);
        load_x(0);
    }

    constexpr uint32_t IDESC1 = idesc_bf16_f32(128, C::N1, false, false);
    constexpr uint32_t IDESC2 = idesc_bf16_f32(128, C::N2, false, false);
    // TMEM lanes 32q..32q+31 are only reachable from warps with warp % 4 == q: a warp serves the (tile, quadrant)
    // tasks of its own quadrant, tiles strided by the number of warps sharing that quadrant
    const int q = warp & 3;
    const int q_warps = (C::WARPS - q + 3) / 4, q_rank = warp >> 2;

#ifdef HSB_IR_PROF
    long long prof_acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long prof_t = clock64();
#endif
    uint32_t it = 0;
    for (int patch = blockIdx.x; patch < p.total; patch += gridDim.x, ++it) {
        const uint32_t par = it & 1;
        HSB_STAMP(9);

        // ---------------- P0: operands of this patch have landed ----------------
        if (p.w_bulk) mbar_wait(bar_w, par); else mbar_wait(bar_tma, par);
        HSB_STAMP(0);
        const int b = coord[par * 4 + 0], pi = coord[par * 4 + 1], pj = coord[par * 4 + 2];
        if (!p.w_bulk) {
            const __nv_bfloat16* src = p.w + (size_t)patch * p.w_row_stride;
            for (int k = tid; k < C::HP; k += C::THREADS) rawW[k] = src[k];
        }
        const bool left = pj == 0, right = pj == p.fw - 1, top = pi == 0, bottom = pi == p.fh - 1;
        auto patch_halo = [&]() {
            if (left || right) {                       // reflect: column -1 <- column 1, column W <- column W-2
                for (int i = tid; i < C::CIN * C::TH; i += C::THREADS) {
                    __nv_bfloat16* row = rawX + (size_t)i * C::TWB + C::XOFF;
                    if (left) row[0] = row[2];
                    if (right) row[C::TW - 1] = row[C::TW - 3];
                }
                __syncthreads();
            }
            if (top || bottom) {
                for (int i = tid; i < C::CIN * C::TW; i += C::THREADS) {
                    int c = i / C::TW, qq = i % C::TW;
                    __nv_bfloat16* ch = rawX + (size_t)c * C::TH * C::TWB + C::XOFF + qq;
                    if (top) ch[0] = ch[2 * C::TWB];
                    if (bottom) ch[(C::TH - 1) * C::TWB] = ch[(C::TH - 3) * C::TWB];
                }
            }
        };
        if (!p.w_bulk) __syncthreads();
        HSB_STAMP(1);

        // ---------------- P1: re-stage into UMMA operand layouts ----------------
        auto restage_x = [&]() {
            // x tile -> A1 (K-major): unit(m, kc) = kc*LBO + (m/8)*SBO + (m%8)*16 bytes = 8 channels of pixel m.
            // Lanes walk consecutive pixels: 2-byte reads of one tile row are contiguous, the 16-byte writes of a warp
            // cover 512 contiguous bytes -> no bank conflicts either way.  Channel CIN is the constant one; k-chunks
            // past it keep whatever finite bytes A2 left there (their B1 columns are zero).
            constexpr int KCX = (C::CIN + 1 + 7) / 8;
            constexpr int CHS = C::TH * C::TWB;                    // elements between channels of the raw tile
            for (int i = tid; i < KCX * C::T; i += C::THREADS) {
                const int m = i % C::T, kc = i / C::T;
                const unsigned short* src = reinterpret_cast<const unsigned short*>(rawX) + (m / C::TW) * C::TWB +
                                            (m % C::TW) + C::XOFF + kc * 8 * CHS;
                uint32_t h[8];
                if (kc < C::CIN / 8) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) h[e] = src[e * CHS];
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        constexpr int k0 = (C::CIN / 8) * 8;
                        h[e] = (k0 + e < C::CIN) ? (uint32_t)src[e * CHS] : ((k0 + e == C::CIN) ? 0x3F80u : 0u);
                    }
                }
                *reinterpret_cast<uint4*>(sm + C::OFF_A1 + kc * C::A1_LBO + (m >> 3) * C::A1_SBO + (m & 7) * 16) =
                    make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16));
            }
        };
        {
            // W1 -> B1 (K-major): unit(n, kc) = kc*LBO + (n/8)*SBO + (n%8)*16; row n = s1[n]*W1[n][:], b1[n] at k=CIN.
            // Every unit of B1 is rewritten (zeros for padding rows / columns): the region is shared with A2.
            for (int i = tid; i < C::N1 * (C::K1 / 8); i += C::THREADS) {
                const int n = i % C::N1, kc = i / C::N1;
                uint32_t v[4] = {0u, 0u, 0u, 0u};
                if (n < C::HID && kc * 8 <= C::CIN) {
                    const float sc = s1[n];
                    const __nv_bfloat16* src = rawW + n * C::CIN;
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float f[2];
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int k = kc * 8 + e * 2 + h;
                            f[h] = k < C::CIN ? __bfloat162float(src[k]) * sc : (k == C::CIN ? b1[n] : 0.f);
                        }
                        v[e] = pack_bf16(f[0], f[1]);
                    }
                }
                *reinterpret_cast<uint4*>(sm + C::OFF_B1 + kc * C::B1_LBO + (n >> 3) * C::B1_SBO + (n & 7) * 16) =
                    make_uint4(v[0], v[1], v[2], v[3]);
            }
            // W3 -> B2 (K-major, own region: padding rows / columns stay zero from the set-up)
            constexpr int KC2 = (C::HID + 1 + 7) / 8;
            for (int i = tid; i < C::COUT * KC2; i += C::THREADS) {
                const int n = i / KC2, kc = i % KC2;
                const float sc = s3[n];
                const __nv_bfloat16* src = rawW + C::R2 + n * C::HID;
                uint32_t v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float f[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int k = kc * 8 + e * 2 + h;
                        f[h] = k < C::HID ? __bfloat162float(src[k]) * sc : (k == C::HID ? b3[n] : 0.f);
                    }
                    v[e] = pack_bf16(f[0], f[1]);
                }
                *reinterpret_cast<uint4*>(sm + C::OFF_B2 + kc * C::B2_LBO + (n >> 3) * C::B2_SBO + (n & 7) * 16) =
                    make_uint4(v[0], v[1], v[2], v[3]);
            }
            // W2 -> packed taps: w2p[tap][cp] = (s2*W2[2cp][tap], s2*W2[2cp+1][tap]); w2p[9][cp] = (b2, b2)
            for (int i = tid; i < 10 * C::NPAIR; i += C::THREADS) {
                const int tap = i / C::NPAIR, cp = i % C::NPAIR;
                float lo, hi;
                if (tap < 9) {
                    lo = __bfloat162float(rawW[C::R1 + (2 * cp) * 9 + tap]) * s2[2 * cp];
                    hi = __bfloat162float(rawW[C::R1 + (2 * cp + 1) * 9 + tap]) * s2[2 * cp + 1];
                } else {
                    lo = b2[2 * cp]; hi = b2[2 * cp + 1];
                }
                w2p[tap * C::HPW + cp] = pack_bf16(lo, hi);
            }
        }
        // the x tile is only needed now: its flight time hid behind the weight re-stage
        if (p.w_bulk) mbar_wait(bar_tma, par);
        patch_halo();
        if (top || bottom) __syncthreads();
        restage_x();
        fence_proxy_async_smem();          // operand writes -> visible to the tensor core (async proxy)
        tc_fence_before_sync();
        __syncthreads();
        HSB_STAMP(2);

        // ---------------- P2: GEMM1; the next patch's weight row starts to stream in ----------------
        const int next = patch + gridDim.x;
        if (warp == 0 && elect_one()) {
            tc_fence_after_sync();
            for (int t = 0; t < C::M1T; ++t) {
#pragma unroll
                for (int s = 0; s < C::K1 / 16; ++s) {
                    const uint64_t da = smem_desc(a1_addr + 2 * s * C::A1_LBO + t * 16 * C::A1_SBO, C::A1_LBO, C::A1_SBO, SWZ_NONE);
                    const uint64_t db = smem_desc(b1_addr + 2 * s * C::B1_LBO, C::B1_LBO, C::B1_SBO, SWZ_NONE);
                    umma_bf16(tmem + t * C::N1, da, db, IDESC1, s > 0);
                }
                umma_commit(bar_mma1 + t);
            }
        }
        if (warp == 1 && elect_one() && next < p.total) announce_and_load_w(next, par ^ 1);     // rawW was consumed in P1

        // ---------------- P3: epilogue 1 (TMEM -> ReLU6 -> hidden tile, which overwrites the raw x tile) ----------------
        HSB_STAMP(3);
        HSB_STAMP(4);
        for (int t = q_rank; t < C::M1T; t += q_warps) {
            if (t * 128 + q * 32 >= C::T) continue;             // warp-uniform: no real pixels in this quadrant
            // quadrant 0 of every tile holds real pixels, so each tile is waited for by some warp before the barrier
            // below: all of GEMM1 is complete (A1/B1 dead) when P4 starts overwriting region Y
            mbar_wait(bar_mma1 + t, par);
            tc_fence_after_sync();
            const int m = t * 128 + q * 32 + lane;
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + t * C::N1;
            unsigned char* hrow = sm + C::OFF_HID + (size_t)m * (C::HPITCH * 2);
            constexpr int FULL = C::HID / 16, REM = C::HID % 16;
            static_assert(REM == 0 || REM == 4 || REM == 8 || REM == 12, "hidden width must be a multiple of 4");
#pragma unroll
            for (int ch = 0; ch < FULL; ch += 2) {          // 32 columns per round trip: two loads in flight, one wait
                uint32_t v0[16], v1[16];
                tmem_ld16(taddr + ch * 16, v0);
                if (ch + 1 < FULL) tmem_ld16(taddr + (ch + 1) * 16, v1);
                tmem_ld_wait();
                uint32_t o[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = pack_relu6(__uint_as_float(v0[2 * e]), __uint_as_float(v0[2 * e + 1]));
                if (m < C::T) {
                    *reinterpret_cast<uint4*>(hrow + ch * 32) = make_uint4(o[0], o[1], o[2], o[3]);
                    *reinterpret_cast<uint4*>(hrow + ch * 32 + 16) = make_uint4(o[4], o[5], o[6], o[7]);
                }
                if (ch + 1 < FULL) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) o[e] = pack_relu6(__uint_as_float(v1[2 * e]), __uint_as_float(v1[2 * e + 1]));
                    if (m < C::T) {
                        *reinterpret_cast<uint4*>(hrow + (ch + 1) * 32) = make_uint4(o[0], o[1], o[2], o[3]);
                        *reinterpret_cast<uint4*>(hrow + (ch + 1) * 32 + 16) = make_uint4(o[4], o[5], o[6], o[7]);
                    }
                }
            }
            if (REM > 0) {
                uint32_t v8[8], v4[4];
                if (REM >= 8) tmem_ld8(taddr + FULL * 16, v8);
                constexpr int c4 = FULL * 16 + (REM >= 8 ? 8 : 0);
                if (REM % 8 == 4) tmem_ld4(taddr + c4, v4);
                tmem_ld_wait();
                if (REM >= 8) {
                    uint32_t o[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) o[e] = pack_relu6(__uint_as_float(v8[2 * e]), __uint_as_float(v8[2 * e + 1]));
                    if (m < C::T) *reinterpret_cast<uint4*>(hrow + FULL * 32) = make_uint4(o[0], o[1], o[2], o[3]);
                }
                if (REM % 8 == 4) {
                    uint32_t o0 = pack_relu6(__uint_as_float(v4[0]), __uint_as_float(v4[1]));
                    uint32_t o1 = pack_relu6(__uint_as_float(v4[2]), __uint_as_float(v4[3]));
                    if (m < C::T) *reinterpret_cast<uint2*>(hrow + c4 * 2) = make_uint2(o0, o1);
                }
            }
        }
        tc_fence_before_sync();
        __syncthreads();
        HSB_STAMP(5);

        // ---------------- P4: depthwise 3x3 + BN2 + ReLU6 -> A2 (which overwrites A1/B1) ----------------
        {
            const uint32_t* hid = reinterpret_cast<const uint32_t*>(sm + C::OFF_HID);
            const __nv_bfloat162 zero = __floats2bfloat162_rn(0.f, 0.f), six = __floats2bfloat162_rn(6.f, 6.f);
            auto column = [&](int cp, int v, unsigned char* dst, int dst_step) {
                __nv_bfloat162 wt[9];
#pragma unroll
                for (int k = 0; k < 9; ++k) wt[k] = *reinterpret_cast<const __nv_bfloat162*>(&w2p[k * C::HPW + cp]);
                const __nv_bfloat162 bias = *reinterpret_cast<const __nv_bfloat162*>(&w2p[9 * C::HPW + cp]);
                const uint32_t* col = hid + v * C::HPW + cp;               // pixel (r, v + kx) at (r*TW + v + kx)*HPW
                __nv_bfloat162 r0[3], r1[3], r2[3];
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    uint32_t a = col[(0 * C::TW + kx) * C::HPW], bq = col[(1 * C::TW + kx) * C::HPW];
                    r0[kx] = *reinterpret_cast<__nv_bfloat162*>(&a);
                    r1[kx] = *reinterpret_cast<__nv_bfloat162*>(&bq);
                }
#pragma unroll
                for (int u = 0; u < C::PH; ++u) {
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        uint32_t a = col[((u + 2) * C::TW + kx) * C::HPW];
                        r2[kx] = *reinterpret_cast<__nv_bfloat162*>(&a);
                    }
                    // three independent partial sums (one per kernel row) keep the FMA chain short
                    __nv_bfloat162 a0 = __hfma2(wt[0], r0[0], bias), a1 = __hmul2(wt[3], r1[0]), a2 = __hmul2(wt[6], r2[0]);
                    a0 = __hfma2(wt[1], r0[1], a0); a1 = __hfma2(wt[4], r1[1], a1); a2 = __hfma2(wt[7], r2[1], a2);
                    a0 = __hfma2(wt[2], r0[2], a0); a1 = __hfma2(wt[5], r1[2], a1); a2 = __hfma2(wt[8], r2[2], a2);
                    __nv_bfloat162 acc = __hadd2(__hadd2(a0, a1), a2);
                    acc = __hmin2(__hmax2(acc, zero), six);
                    *reinterpret_cast<uint32_t*>(dst + u * dst_step) = *reinterpret_cast<uint32_t*>(&acc);
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) { r0[kx] = r1[kx]; r1[kx] = r2[kx]; }
                }
            };
            // two columns per thread, walked in lockstep: twice the independent work per warp
            auto column_pair = [&](int cp, int va, int vb, unsigned char* dsta, unsigned char* dstb, int dst_step) {
                __nv_bfloat162 wt[9];
#pragma unroll
                for (int k = 0; k < 9; ++k) wt[k] = *reinterpret_cast<const __nv_bfloat162*>(&w2p[k * C::HPW + cp]);
                const __nv_bfloat162 bias = *reinterpret_cast<const __nv_bfloat162*>(&w2p[9 * C::HPW + cp]);
                const uint32_t* cola = hid + va * C::HPW + cp;
                const uint32_t* colb = hid + vb * C::HPW + cp;
                __nv_bfloat162 a0r[3], a1r[3], a2r[3], b0r[3], b1r[3], b2r[3];
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    uint32_t t0 = cola[(0 * C::TW + kx) * C::HPW], t1 = cola[(1 * C::TW + kx) * C::HPW];
                    uint32_t t2 = colb[(0 * C::TW + kx) * C::HPW], t3 = colb[(1 * C::TW + kx) * C::HPW];
                    a0r[kx] = *reinterpret_cast<__nv_bfloat162*>(&t0); a1r[kx] = *reinterpret_cast<__nv_bfloat162*>(&t1);
                    b0r[kx] = *reinterpret_cast<__nv_bfloat162*>(&t2); b1r[kx] = *reinterpret_cast<__nv_bfloat162*>(&t3);
                }
#pragma unroll
                for (int u = 0; u < C::PH; ++u) {
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        uint32_t t0 = cola[((u + 2) * C::TW + kx) * C::HPW], t1 = colb[((u + 2) * C::TW + kx) * C::HPW];
                        a2r[kx] = *reinterpret_cast<__nv_bfloat162*>(&t0); b2r[kx] = *reinterpret_cast<__nv_bfloat162*>(&t1);
                    }
                    __nv_bfloat162 pa0 = __hfma2(wt[0], a0r[0], bias), pa1 = __hmul2(wt[3], a1r[0]), pa2 = __hmul2(wt[6], a2r[0]);
                    __nv_bfloat162 pb0 = __hfma2(wt[0], b0r[0], bias), pb1 = __hmul2(wt[3], b1r[0]), pb2 = __hmul2(wt[6], b2r[0]);
                    pa0 = __hfma2(wt[1], a0r[1], pa0); pa1 = __hfma2(wt[4], a1r[1], pa1); pa2 = __hfma2(wt[7], a2r[1], pa2);
                    pb0 = __hfma2(wt[1], b0r[1], pb0); pb1 = __hfma2(wt[4], b1r[1], pb1); pb2 = __hfma2(wt[7], b2r[1], pb2);
                    pa0 = __hfma2(wt[2], a0r[2], pa0); pa1 = __hfma2(wt[5], a1r[2], pa1); pa2 = __hfma2(wt[8], a2r[2], pa2);
                    pb0 = __hfma2(wt[2], b0r[2], pb0); pb1 = __hfma2(wt[5], b1r[2], pb1); pb2 = __hfma2(wt[8], b2r[2], pb2);
                    __nv_bfloat162 ra = __hmin2(__hmax2(__hadd2(__hadd2(pa0, pa1), pa2), zero), six);
                    __nv_bfloat162 rb = __hmin2(__hmax2(__hadd2(__hadd2(pb0, pb1), pb2), zero), six);
                    *reinterpret_cast<uint32_t*>(dsta + u * dst_step) = *reinterpret_cast<uint32_t*>(&ra);
                    *reinterpret_cast<uint32_t*>(dstb + u * dst_step) = *reinterpret_cast<uint32_t*>(&rb);
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) { a0r[kx] = a1r[kx]; a1r[kx] = a2r[kx]; b0r[kx] = b1r[kx]; b1r[kx] = b2r[kx]; }
                }
            };
            if (warp < C::DWW) {
                if (lane < C::MAINP) {
                    const int cp = lane;
                    // pixel m = u*PW + v: 128-byte swizzled row m, 16-byte chunk (cp/4) ^ (m%8); m%8 == v%8 for all u
                    auto a2_dst = [&](int v) { return sm + C::OFF_A2 + v * 128 + ((((cp >> 2) ^ (v & 7)) << 4) | ((cp & 3) << 2)); };
                    if (C::PW == 2 * C::DWW) {
                        column_pair(cp, warp, warp + C::DWW, a2_dst(warp), a2_dst(warp + C::DWW), C::PW * 128);
                    } else {
#pragma unroll 1
                        for (int v = warp; v < C::PW; v += C::DWW) column(cp, v, a2_dst(v), C::PW * 128);
                    }
                }
            } else if (C::TAILP > 0) {
                // tail warp: channel pairs 32.. (K >= 64, non-swizzled units) for every column, then the constant-one column
                const int v = lane >> 1;
                if ((lane & 1) < C::TAILP && v < C::PW) {
                    const int cp = 32 + (lane & 1);
                    unsigned char* dst = sm + C::OFF_A2T + (v >> 3) * C::A2T_SBO + (v & 7) * 16 + (cp - 32) * 4;
                    column(cp, v, dst, (C::PW / 8) * C::A2T_SBO);
                }
            }
            // K columns HID..K2-1 of A2: the constant one that carries BN3's shift, then zeros, written as whole / half
            // 16-byte units.  Rewritten every patch because region Y also hosts A1/B1 (whose bytes must never meet the
            // tensor core as padding).  HID is a multiple of 4, so the first unit is either whole or its upper half.
            {
                constexpr int U0 = C::HID / 8, U1 = C::K2 / 8;              // 16-byte units [U0, U1) contain padding
                for (int i = tid; i < C::M2T * 128 * (U1 - U0); i += C::THREADS) {
                    const int m = i / (U1 - U0), u = U0 + i % (U1 - U0);
                    unsigned char* dst;
                    if (u < 8) dst = sm + C::OFF_A2 + m * 128 + ((u ^ (m & 7)) << 4);
                    else dst = sm + C::OFF_A2T + (u - 8) * C::A2T_LBO + (m >> 3) * C::A2T_SBO + (m & 7) * 16;
                    if (u == U0) {
                        if (C::HID % 8 == 0) *reinterpret_cast<uint4*>(dst) = make_uint4(0x00003F80u, 0u, 0u, 0u);
                        else *reinterpret_cast<uint2*>(dst + 8) = make_uint2(0x00003F80u, 0u);
                    } else {
                        *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
                    }
                }
            }
        }
        fence_proxy_async_smem();
        tc_fence_before_sync();
        __syncthreads();
        HSB_STAMP(6);

        // ---------------- P5: GEMM2 (accumulators reuse GEMM1's TMEM columns) ----------------
        // region X is free: the depthwise phase has consumed the hidden tile
        if (warp == 1 && elect_one() && next < p.total) load_x(par ^ 1);
        if (warp == 0 && elect_one()) {
            tc_fence_after_sync();
            for (int t = 0; t < C::M2T; ++t) {
#pragma unroll
                for (int s = 0; s < C::K2 / 16; ++s) {
                    uint64_t da;
                    if (s < 4) da = smem_desc(a2_addr + t * 128 * 128 + s * 32, 16, 1024, SWZ_128B);
                    else da = smem_desc(a2t_addr + 2 * (s - 4) * C::A2T_LBO + t * 16 * C::A2T_SBO, C::A2T_LBO, C::A2T_SBO, SWZ_NONE);
                    const uint64_t db = smem_desc(b2_addr + 2 * s * C::B2_LBO, C::B2_LBO, C::B2_SBO, SWZ_NONE);
                    umma_bf16(tmem + t * C::N2, da, db, IDESC2, s > 0);
                }
                umma_commit(bar_mma2 + t);
            }
        }

        HSB_STAMP(7);
        // ---------------- P6: epilogue 2 (TMEM -> bf16 -> NCHW) ----------------
        for (int t = q_rank; t < C::M2T; t += q_warps) {
            mbar_wait(bar_mma2 + t, par);
            tc_fence_after_sync();
            const int m = t * 128 + q * 32 + lane;
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + t * C::N2;
            const int u = m / C::PW, vv = m % C::PW;
            __nv_bfloat16* yp = p.y + (((size_t)b * C::COUT) * p.H + (size_t)pi * C::PH + u) * p.W + (size_t)pj * C::PW + vv;
            const size_t plane = (size_t)p.H * p.W;
#pragma unroll
            for (int c0 = 0; c0 < C::COUT; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(taddr + c0, v);          // N2 is a multiple of 16, so this never leaves the accumulator
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 16; ++e)
                    if (c0 + e < C::COUT && m < C::O) yp[(size_t)(c0 + e) * plane] = __float2bfloat16_rn(__uint_as_float(v[e]));
            }
        }
        tc_fence_before_sync();
        __syncthreads();       // TMEM and the operand buffers are reused by the next patch
        tc_fence_after_sync();
        HSB_STAMP(8);
    }
#ifdef HSB_IR_PROF
    if (tid == 0) {
        for (int k = 0; k < 10; ++k) p.prof[(size_t)blockIdx.x * 12 + k] = prof_acc[k];
        p.prof[(size_t)blockIdx.x * 12 + 10] = it;
    }
#endif

    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, C::TMEM_COLS);
}

// ---- host side ------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static std::once_flag once;
    static EncodeTiledFn fn = nullptr;
    std::call_once(once, [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
        else
            cudaGetLastError();
    });
    return fn;
}

template <class C>
static int launch_tc(const void* x, const IRTCParams& p, cudaStream_t st) {
    EncodeTiledFn encode = get_encode_fn();
    if (!encode) return fail(HSB_ERR_CUDA, "patch_ir_tc: cuTensorMapEncodeTiled is not available from the driver");
    CUtensorMap map;
    const cuuint64_t dims[4] = {(cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)C::CIN, (cuuint64_t)p.B};
    const cuuint64_t strides[3] = {(cuuint64_t)p.W * 2, (cuuint64_t)p.W * p.H * 2, (cuuint64_t)p.W * p.H * C::CIN * 2};
    const cuuint32_t box[4] = {(cuuint32_t)C::TWB, (cuuint32_t)C::TH, (cuuint32_t)C::CIN, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(HSB_ERR_CUDA, "patch_ir_tc: cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    auto kern = patch_ir_tc_kernel<C>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return fail(HSB_ERR_CUDA, std::string("patch_ir_tc attr: ") + cudaGetErrorString(e));
    // ask for the largest shared-memory carve-out, otherwise the driver may size it for a single CTA
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    static const int force_ctas = [] { const char* v = getenv("HSB_IR_CTAS"); return v ? atoi(v) : 0; }();
    const int ctas = force_ctas > 0 ? force_ctas : C::CTAS;
    static const bool verbose = [] { const char* v = getenv("HSB_VERBOSE"); return v && v[0] == '1'; }();
    if (verbose) {
        int resident = -1;
        cudaError_t oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, C::THREADS, (size_t)C::SMEM_BYTES);
        fprintf(stderr, "[hsb] patch_ir_tc<%d,%d,%d,%d>: %d threads, %d B smem, occupancy query -> %d (%s), launching %d CTAs/SM\n",
                C::CIN, C::HID, C::COUT, C::PH, C::THREADS, C::SMEM_BYTES, resident, cudaGetErrorString(oe), ctas);
        cudaGetLastError();
    }
    const int grid = std::min(p.total, std::max(1, device_sm_count()) * ctas);
#ifdef HSB_IR_PROF
    {   // profiling build: per-phase cycle sums of thread 0 of every CTA, printed after a synchronising launch
        static long long* dprof = nullptr;
        if (!dprof) cudaMalloc(&dprof, 4096 * 12 * sizeof(long long));
        cudaMemsetAsync(dprof, 0, 4096 * 12 * sizeof(long long), st);
        IRTCParams pp = p;
        pp.prof = dprof;
        kern<<<grid, C::THREADS, C::SMEM_BYTES, st>>>(map, pp);
        cudaStreamSynchronize(st);
        static long long host[4096 * 12];
        cudaMemcpy(host, dprof, sizeof(long long) * grid * 12, cudaMemcpyDeviceToHost);
        double sum[10] = {0}; double patches = 0;
        for (int g = 0; g < grid; ++g) { for (int k = 0; k < 10; ++k) sum[k] += (double)host[g * 12 + k]; patches += (double)host[g * 12 + 10]; }
        static const char* names[10] = {"P0 wait loads", "P0 halo patch", "P1 restage", "P2 issue GEMM1", "P3 wait GEMM1", "P3 epilogue1",
                                        "P4 depthwise", "P5 issue GEMM2", "P6 epilogue2", "loop top"};
        double tot = 0; for (int k = 0; k < 10; ++k) tot += sum[k];
        fprintf(stderr, "[hsb-prof] patch_ir_tc<%d,%d,%d,%d> grid %d, %.0f patches, %.0f cycles/patch (thread 0)\n", C::CIN, C::HID, C::COUT, C::PH,
                grid, patches, tot / patches);
        for (int k = 0; k < 10; ++k) fprintf(stderr, "[hsb-prof]   %-16s %8.0f cycles/patch  %5.1f %%\n", names[k], sum[k] / patches, 100.0 * sum[k] / tot);
        return check_launch("patch_ir_tc launch");
    }
#endif
    kern<<<grid, C::THREADS, C::SMEM_BYTES, st>>>(map, p);
    note_kernel("patch_ir_tc_kernel");
    return check_launch("patch_ir_tc launch");
}

int launch_patch_ir_tc(const void* x, const void* w, void* y, const float* const* bn, int B, int Cin, int hid, int Cout,
                       int H, int W, int fh, int fw, int residual, int64_t w_row_stride, cudaStream_t st, bool* handled) {
    *handled = false;
    static const bool disabled = [] { const char* e = getenv("HSB_DISABLE_TC"); return e && e[0] == '1'; }();
    if (disabled || residual) return HSB_OK;
    const int ps = H / fh;
    if (W / fw != ps || (ps != 16 && ps != 8)) return HSB_OK;
    if ((W % 8) != 0 || (reinterpret_cast<uintptr_t>(x) & 15)) return HSB_OK;      // TMA: 16-byte strides / base
    IRTCParams p;
    p.w = reinterpret_cast<const __nv_bfloat16*>(w);
    p.y = reinterpret_cast<__nv_bfloat16*>(y);
    for (int i = 0; i < 6; ++i) p.bn[i] = bn[i];
    p.B = B; p.H = H; p.W = W; p.fh = fh; p.fw = fw; p.w_row_stride = w_row_stride; p.total = B * fh * fw;
    const int64_t hp = (int64_t)Cin * hid + 9 * hid + (int64_t)hid * Cout;
    // bulk copies move ceil16(hp*2) bytes per row: the row (incl. that round-up) must stay inside its stride
    p.w_bulk = ((reinterpret_cast<uintptr_t>(w) & 15) == 0) && ((w_row_stride * 2) % 16 == 0) &&
               (((hp * 2 + 15) / 16 * 16) <= w_row_stride * 2);
#define HSB_TC_CASE(CI, HD, CO, PS)                                           \
    if (Cin == CI && hid == HD && Cout == CO && ps == PS) {                   \
        *handled = true;                                                      \
        return launch_tc<IRTC<CI, HD, CO, PS>>(x, p, st);                     \
    }
    HSB_TC_CASE(34, 68, 19, 16)     // HyperSeg-M level 4
    HSB_TC_CASE(26, 52, 19, 16)     // HyperSeg-S Cityscapes level 4
    HSB_TC_CASE(22, 44, 12, 16)     // HyperSeg-S CamVid level 4
    HSB_TC_CASE(24, 48, 16, 8)      // HyperSeg-M / CamVid level 3
    HSB_TC_CASE(14, 28, 8, 8)       // HyperSeg-S Cityscapes level 3
#undef HSB_TC_CASE
    return HSB_OK;
}

}  // namespace hsb
