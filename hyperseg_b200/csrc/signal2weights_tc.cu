// Weight head on the tensor cores (tcgen05): signal -> per-patch weights, bf16, patch-major output.
//
// Same arithmetic as signal2weights.cu (reference hyperseg/models/hyperseg_v1_0.py:315-326): per group g
//     Wout[n, o] = sum_k S[n, idx + g*K + k] * Ws[o, k]        n = (b, i, j) position, K = sig_ch / groups
// i.e. a GEMM with a tiny K (12..80) whose cost is writing Wout (16 MB per 512x1024 frame).  Organisation:
//
//   * the static head weights are packed ONCE (hsb_head_pack) into the UMMA K-major operand layout, zero padded
//     to K % 16 == 0 and to 256-channel tiles, so a tile of B is one cp.async.bulk;
//   * a CTA owns 128 consecutive positions and a contiguous range of (group, 256-channel tile) items: the producer
//     warp copies the item's signal slab [K x 128] from NCHW straight into the MN-major operand layout (16-byte
//     units of 8 positions -- no transposition), streams the packed B tile (2-stage operand ring) and issues
//     128x256xK MMAs into a 2-stage TMEM accumulator; eight epilogue warps drain TMEM -> bf16 -> a shared staging
//     tile -> coalesced row stores, one item behind the MMAs.
#include "common.cuh"
#include "tcgen05.cuh"

namespace hsb {

constexpr int HD_NT = 256;                 // output channels per tile
constexpr int HD_M = 128;                  // positions per CTA
constexpr int HD_EPI_WARPS = 8;            // 4 TMEM quadrants x 2 column halves
constexpr int HD_THREADS = 32 * (HD_EPI_WARPS + 1);     // + 1 producer / MMA warp
constexpr int HD_STAGE_PITCH = HD_NT * 2 + 16;

struct HeadTCParams {
    const __nv_bfloat16* s;
    const __nv_bfloat16* packed;      // [groups][otiles][Kpad/8][NT/8][8][8]
    __nv_bfloat16* out;
    int NTOT;                         // B * P positions
    int P;
    int sig_index, spg, kpad, opg, hp, groups, otiles;
    int items;                        // groups * otiles work items per position tile
    int splits;                       // CTAs sharing one position tile
    int64_t ssb, ssc;                 // signal strides (elements); position stride is 1
    int64_t row_stride;               // output row stride (elements)
};

__host__ __device__ inline size_t head_smem_bytes(int kpad) {
    size_t a = 2 * (size_t)kpad * HD_M * 2;             // two A stages (signal slab of the item's group)
    size_t b = 2 * (size_t)HD_NT * kpad * 2;            // two B stages
    size_t st = (size_t)HD_M * HD_STAGE_PITCH;          // output staging
    return a + b + st + 128 + 1024;
}

// A CTA owns 128 positions and a contiguous range of (group, channel-tile) items.  Warp 8 is the producer: it copies
// the item's signal slab into the MN-major A operand (all 32 lanes), streams the packed B tile with cp.async.bulk and
// issues the MMAs (lane 0) into a two-stage TMEM accumulator.  Warps 0-7 drain: quadrant = warp % 4, column half =
// warp / 4; TMEM -> bf16 -> staging rows -> coalesced stores, one item behind the MMAs.
__global__ void __launch_bounds__(HD_THREADS, 1) signal2weights_tc_kernel(const HeadTCParams p) {
    extern __shared__ unsigned char smem_dyn[];
    // align by pointer arithmetic on the __shared__ array so the compiler keeps the address space (LDS/STS, not generic)
    unsigned char* sm = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int kpad = p.kpad;
    const size_t a_bytes = (size_t)kpad * HD_M * 2, b_bytes = (size_t)HD_NT * kpad * 2;
    unsigned char* a_sm = sm;
    unsigned char* b_sm = a_sm + 2 * a_bytes;
    unsigned char* st_sm = b_sm + 2 * b_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(st_sm + (size_t)HD_M * HD_STAGE_PITCH);
    uint64_t* b_full = bars;          // [2] B tile landed
    uint64_t* s_empty = bars + 2;     // [2] operand stage (A and B) consumed by the tensor core
    uint64_t* d_full = bars + 4;      // [2] accumulator ready
    uint64_t* d_empty = bars + 6;     // [2] accumulator drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

    const int n0 = blockIdx.x * HD_M;
    const int per = (p.items + p.splits - 1) / p.splits;
    const int it0 = blockIdx.y * per, it1 = min(p.items, it0 + per);
    const int nitems = it1 - it0;
    if (nitems <= 0) return;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(b_full + i, 1);
            mbar_init(s_empty + i, 1);
            mbar_init(d_full + i, 1);
            mbar_init(d_empty + i, HD_EPI_WARPS);
        }
        mbar_fence_init();
    }
    if (warp == HD_EPI_WARPS) tmem_alloc(tmem_slot, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const int a_lbo = (HD_M / 8) * 128, b_lbo = (HD_NT / 8) * 128;

    if (warp == HD_EPI_WARPS) {
        // ================= producer / MMA warp =================
        const uint32_t idesc = idesc_bf16_f32(HD_M, HD_NT, /*A MN-major*/ true, false);
        int stage_group0 = -1, stage_group1 = -1;
        auto stage_operands = [&](int j) {          // whole warp: A slab (if the group changed) + B tile of item j
            const int st = j & 1, it = it0 + j, g = it / p.otiles, t = it % p.otiles;
            if (j >= 2) mbar_wait(s_empty + st, ((j >> 1) - 1) & 1);      // MMAs of item j-2 done with this stage
            if ((st ? stage_group1 : stage_group0) != g) {
                unsigned char* a_dst = a_sm + st * a_bytes;
                // unit(mc, k) = (k/8)*LBO + mc*128 + (k%8)*16: 8 consecutive positions of signal channel k
                const int units = kpad * (HD_M / 8);
                for (int base = 0; base < units; base += 32 * 8) {      // 8 independent 16-byte loads in flight per lane
                    uint4 v[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int i = base + e * 32 + lane;
                        const int mc = i % (HD_M / 8), k = i / (HD_M / 8);
                        const int n = n0 + mc * 8;
                        v[e] = make_uint4(0, 0, 0, 0);
                        if (i < units && k < p.spg && n < p.NTOT) {
                            const int b = n / p.P, pp = n % p.P;          // P % 8 == 0: a unit never straddles images
                            v[e] = __ldg(reinterpret_cast<const uint4*>(p.s + (size_t)b * p.ssb +
                                                                        (size_t)(p.sig_index + g * p.spg + k) * p.ssc + pp));
                        }
                    }
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int i = base + e * 32 + lane;
                        const int mc = i % (HD_M / 8), k = i / (HD_M / 8);
                        if (i < units) *reinterpret_cast<uint4*>(a_dst + (k >> 3) * a_lbo + mc * 128 + (k & 7) * 16) = v[e];
                    }
                }
                if (st) stage_group1 = g; else stage_group0 = g;
                fence_proxy_async_smem();
            }
            __syncwarp();
            if (elect_one()) {
                mbar_arrive_expect_tx(b_full + st, (uint32_t)b_bytes);
                bulk_g2s(b_sm + st * b_bytes, p.packed + ((size_t)g * p.otiles + t) * HD_NT * kpad, (uint32_t)b_bytes,
                         b_full + st);
            }
        };
        stage_operands(0);
        for (int j = 0; j < nitems; ++j) {
            const int st = j & 1;
            if (j + 1 < nitems) stage_operands(j + 1);          // overlaps the MMAs / epilogue of item j
            if (elect_one()) {
                mbar_wait(b_full + st, (j >> 1) & 1);
                if (j >= 2) mbar_wait(d_empty + st, ((j >> 1) - 1) & 1);      // epilogue drained this accumulator
                tc_fence_after_sync();
                const uint32_t a_addr = smem_u32(a_sm + st * a_bytes), b_addr = smem_u32(b_sm + st * b_bytes);
                for (int s = 0; s < kpad / 16; ++s) {
                    const uint64_t da = smem_desc(a_addr + 2 * s * a_lbo, a_lbo, 128, SWZ_NONE);
                    const uint64_t db = smem_desc(b_addr + 2 * s * b_lbo, b_lbo, 128, SWZ_NONE);
                    umma_bf16(tmem + st * HD_NT, da, db, idesc, s > 0);
                }
                umma_commit(d_full + st);       // accumulator ready
                umma_commit(s_empty + st);      // operand stage reusable
            }
            __syncwarp();
        }
    } else {
        // ================= epilogue warps =================
        const int q = warp & 3, half = warp >> 2;
        const int row = q * 32 + lane;
        constexpr int HC = HD_NT / 2;                     // columns per warp
        unsigned char* my_row = st_sm + (size_t)row * HD_STAGE_PITCH + half * HC * 2;
        for (int j = 0; j < nitems; ++j) {
            const int st = j & 1, it = it0 + j, g = it / p.otiles, t = it % p.otiles;
            const int o_base = g * p.opg + t * HD_NT;
            const int o_end = min(min((g + 1) * p.opg, p.hp), o_base + HD_NT);
            const int nvalid = min(HC, max(o_end - o_base, 0) - half * HC);   // valid columns in this warp's half (may be <= 0)
            const int ncols = min(HC, (max(nvalid, 0) + 31) & ~31);
            mbar_wait(d_full + st, (j >> 1) & 1);
            tc_fence_after_sync();
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + st * HD_NT + half * HC;
            for (int c = 0; c < ncols; c += 32) {
                uint32_t v0[16], v1[16];
                tmem_ld16(taddr + c, v0);
                tmem_ld16(taddr + c + 16, v1);
                tmem_ld_wait();
                uint32_t o[16];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    __nv_bfloat162 a = __floats2bfloat162_rn(__uint_as_float(v0[2 * e]), __uint_as_float(v0[2 * e + 1]));
                    __nv_bfloat162 bq = __floats2bfloat162_rn(__uint_as_float(v1[2 * e]), __uint_as_float(v1[2 * e + 1]));
                    o[e] = *reinterpret_cast<uint32_t*>(&a);
                    o[8 + e] = *reinterpret_cast<uint32_t*>(&bq);
                }
                uint4* dst = reinterpret_cast<uint4*>(my_row + c * 2);
                dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
                dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
                dst[2] = make_uint4(o[8], o[9], o[10], o[11]);
                dst[3] = make_uint4(o[12], o[13], o[14], o[15]);
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(d_empty + st);          // TMEM stage free again
            // coalesced write-out of this warp's 32 rows x its column half
            if (nvalid > 0) {
                const int ob = o_base + half * HC;
                const bool pairs = (ob & 1) == 0;
                const int nrows = min(32, p.NTOT - (n0 + q * 32));
                const unsigned char* sbase = st_sm + (size_t)(q * 32) * HD_STAGE_PITCH + half * HC * 2;
                __nv_bfloat16* dbase = p.out + (size_t)(n0 + q * 32) * p.row_stride + ob;
                if (pairs) {
                    // lane = column pair, rows unrolled: 8 independent shared loads / global stores in flight;
                    // for a fixed row the 32 lanes write 128 contiguous bytes
                    for (int c = lane * 2; c < nvalid; c += 64) {
                        const bool two = c + 1 < nvalid;
#pragma unroll 8
                        for (int r = 0; r < 32; ++r) {
                            if (r < nrows) {
                                const unsigned char* sp = sbase + (size_t)r * HD_STAGE_PITCH + c * 2;
                                __nv_bfloat16* dp = dbase + (size_t)r * p.row_stride + c;
                                if (two) *reinterpret_cast<uint32_t*>(dp) = *reinterpret_cast<const uint32_t*>(sp);
                                else *dp = *reinterpret_cast<const __nv_bfloat16*>(sp);
                            }
                        }
                    }
                } else {
                    for (int c = lane; c < nvalid; c += 32) {
#pragma unroll 8
                        for (int r = 0; r < 32; ++r)
                            if (r < nrows)
                                dbase[(size_t)r * p.row_stride + c] =
                                    *reinterpret_cast<const __nv_bfloat16*>(sbase + (size_t)r * HD_STAGE_PITCH + c * 2);
                    }
                }
            }
            __syncwarp();       // staging rows are rewritten by the next item
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == HD_EPI_WARPS) tmem_dealloc(tmem, 512);
}

// ---- packing of the static head weights into the UMMA operand layout ---------------------------------------------------
template <typename T>
__global__ void head_pack_kernel(const T* __restrict__ ws, __nv_bfloat16* __restrict__ out, const float* __restrict__ scale,
                                 int out_ch, int spg, int kpad, int opg, int groups, int otiles) {
    const size_t total = (size_t)groups * otiles * HD_NT * kpad;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        // destination index -> (g, t, kc, nc, n8, k8)
        size_t r = i;
        const int k8 = r % 8; r /= 8;
        const int n8 = r % 8; r /= 8;
        const int nc = r % (HD_NT / 8); r /= (HD_NT / 8);
        const int kc = r % (kpad / 8); r /= (kpad / 8);
        const int t = r % otiles; const int g = r / otiles;
        const int ol = t * HD_NT + nc * 8 + n8, k = kc * 8 + k8;
        float v = 0.f;
        if (ol < opg && k < spg) {
            const int o = g * opg + ol;
            v = ld_f(ws + (size_t)o * spg + k);
            if (scale) v *= scale[o];
        }
        out[i] = __float2bfloat16_rn(v);
    }
}

}  // namespace hsb

using namespace hsb;

extern "C" int64_t hsb_head_packed_elems(int sig_ch, int out_ch, int groups) {
    if (sig_ch <= 0 || out_ch <= 0 || groups <= 0 || sig_ch % groups || out_ch % groups) return -1;
    const int spg = sig_ch / groups, opg = out_ch / groups;
    const int kpad = (spg + 15) / 16 * 16, otiles = ceil_div(opg, HD_NT);
    return (int64_t)groups * otiles * HD_NT * kpad;
}

extern "C" int hsb_head_pack(const void* ws, void* packed, const float* row_scale, int sig_ch, int out_ch, int groups,
                             int dtype, void* stream) {
    HSB_REQUIRE(ws && packed, HSB_ERR_INVALID_ARG, "head_pack: null pointer");
    HSB_REQUIRE(hsb_head_packed_elems(sig_ch, out_ch, groups) > 0, HSB_ERR_INVALID_ARG, "head_pack: bad dimensions");
    HSB_REQUIRE(dtype == HSB_F32 || dtype == HSB_BF16, HSB_ERR_INVALID_ARG, "head_pack: bad dtype");
    const int spg = sig_ch / groups, opg = out_ch / groups;
    const int kpad = (spg + 15) / 16 * 16, otiles = ceil_div(opg, HD_NT);
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = std::max(1, device_sm_count()) * 4;
    if (dtype == HSB_F32)
        head_pack_kernel<float><<<blocks, 256, 0, st>>>((const float*)ws, (__nv_bfloat16*)packed, row_scale, out_ch, spg,
                                                        kpad, opg, groups, otiles);
    else
        head_pack_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)ws, (__nv_bfloat16*)packed, row_scale,
                                                                out_ch, spg, kpad, opg, groups, otiles);
    return check_launch("head_pack launch");
}

extern "C" int hsb_signal2weights_packed_fwd(const void* s, const void* packed, void* w_out,
                                             int B, int sig_index, int sig_ch, int out_ch, int hp, int groups,
                                             int fh, int fw, int64_t s_stride_b, int64_t s_stride_c,
                                             int64_t out_row_stride, void* stream) {
    HSB_REQUIRE(s && packed && w_out, HSB_ERR_INVALID_ARG, "signal2weights_packed: null pointer");
    HSB_REQUIRE(B > 0 && sig_ch > 0 && out_ch > 0 && hp > 0 && groups > 0 && fh > 0 && fw > 0 && sig_index >= 0,
                HSB_ERR_INVALID_ARG, "signal2weights_packed: bad dimension");
    HSB_REQUIRE(sig_ch % groups == 0 && out_ch % groups == 0 && hp <= out_ch, HSB_ERR_INVALID_ARG,
                "signal2weights_packed: channels not divisible by groups");
    const int P = fh * fw;
    HSB_REQUIRE(P % 8 == 0, HSB_ERR_UNSUPPORTED, "signal2weights_packed: fh*fw must be a multiple of 8");
    HSB_REQUIRE(out_row_stride >= hp, HSB_ERR_INVALID_ARG, "signal2weights_packed: out_row_stride < hp");
    HSB_REQUIRE(((uintptr_t)s % 16) == 0 && (s_stride_b % 8) == 0 && (s_stride_c % 8) == 0 && ((uintptr_t)packed % 16) == 0,
                HSB_ERR_UNSUPPORTED, "signal2weights_packed: signal / packed weights must be 16-byte aligned");
    HeadTCParams p;
    p.s = (const __nv_bfloat16*)s; p.packed = (const __nv_bfloat16*)packed; p.out = (__nv_bfloat16*)w_out;
    p.NTOT = B * P; p.P = P; p.sig_index = sig_index; p.spg = sig_ch / groups; p.kpad = (p.spg + 15) / 16 * 16;
    p.opg = out_ch / groups; p.hp = hp; p.groups = groups; p.otiles = ceil_div(p.opg, HD_NT);
    p.ssb = s_stride_b; p.ssc = s_stride_c; p.row_stride = out_row_stride;
    const size_t smem = head_smem_bytes(p.kpad);
    HSB_REQUIRE(smem <= 227 * 1024, HSB_ERR_UNSUPPORTED, "signal2weights_packed: sig_ch / groups too large for one CTA");
    cudaError_t e = cudaFuncSetAttribute(signal2weights_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(HSB_ERR_CUDA, std::string("signal2weights_packed attr: ") + cudaGetErrorString(e));
    p.items = groups * p.otiles;
    const int tiles = ceil_div(p.NTOT, HD_M);
    // split the items of a position tile over `splits` CTAs: minimise (waves of CTAs) x (items per CTA + set-up), where
    // the per-CTA set-up (launch, TMEM allocation, barrier init, pipeline fill) is worth about four items (measured:
    // one-item CTAs take ~10 us each)
    const int sms = std::max(1, device_sm_count());
    int splits = 1, best = ceil_div(tiles, sms) * (p.items + 4);
    for (int sp = 2; sp <= p.items; ++sp) {
        const int cost = ceil_div(tiles * sp, sms) * (ceil_div(p.items, sp) + 4);
        if (cost < best) { best = cost; splits = sp; }
    }
    p.splits = splits;
    HSB_REQUIRE(splits <= 65535, HSB_ERR_UNSUPPORTED, "signal2weights_packed: grid too large");
    dim3 grid(tiles, splits);
    signal2weights_tc_kernel<<<grid, HD_THREADS, smem, (cudaStream_t)stream>>>(p);
    return check_launch("signal2weights_packed launch");
}
