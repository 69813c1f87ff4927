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
//
// The same kernel also emits "arranged" rows for the fused inverted-residual MetaBlock (ir_arranged.cuh): the static
// weights are then packed in the block's operand order (hsb_head_pack_arranged), which interleaves the head's groups
// inside a tile -- an item carries its own range of signal channels (all groups its columns use, zero blocks where a
// column belongs to another group) and the BatchNorm scales of the block are folded into the packed rows.
#include <vector>

#include "common.cuh"
#include "ir_arranged.cuh"
#include "tcgen05.cuh"

namespace hsb {

void note_kernel(const char* name);

constexpr int HD_NT = 256;                 // output channels per tile
constexpr int HD_M = 128;                  // positions per CTA
constexpr int HD_EPI_WARPS = 8;            // 4 TMEM quadrants x 2 column halves
constexpr int HD_THREADS = 32 * (HD_EPI_WARPS + 1);     // + 1 producer / MMA warp
constexpr int HD_STAGE_PITCH = HD_NT * 2 + 16;
constexpr int HD_TBL_ITEMS = 64;           // item table entries of one CTA kept in shared memory
constexpr int HD_TBL_BYTES = HD_TBL_ITEMS * 16 + (HD_M / 8) * 8;     // + first-element offsets of the 16 position units

struct HeadTCParams {
    const __nv_bfloat16* s;
    const __nv_bfloat16* packed;      // [groups][otiles][Kpad/8][NT/8][8][8]
    __nv_bfloat16* out;
    int NTOT;                         // B * P positions
    int P;
    int sig_index, spg, kpad, opg, hp, groups, otiles;
    int items;                        // groups * otiles work items per position tile
    int splits;                       // CTAs sharing one position tile
    const int4* table;                // arranged mode: per item {first signal channel, kpad, first output column, columns};
                                      // the packed tile of item i starts at element i * nt * kpad_max
    int kpad_max, sig_end;            // arranged mode: operand stages are sized for kpad_max; signal channels >= sig_end read as 0
    int sig_first, tbl_base;          // first channel of the resident slab; added to the table's first-channel entries
    int64_t ssb, ssc;                 // signal strides (elements); position stride is 1
    int64_t row_stride;               // output row stride (elements)
    long long* prof;                  // profiling build (-DHSB_HEAD_PROF): 8 counters per CTA
};

__host__ __device__ inline size_t head_smem_bytes(int kpad, int nt = HD_NT) {
    size_t a = 2 * (size_t)kpad * HD_M * 2;             // two A stages (signal slab of the item's group)
    size_t b = 2 * (size_t)nt * kpad * 2;               // two B stages
    size_t st = (size_t)HD_M * (nt * 2 + 16);           // output staging
    return a + b + st + 128 + HD_TBL_BYTES + 1024;
}

// arranged mode: the whole signal slice of the head stays resident (one MN-major slab, k-chunks addressed per item)
__host__ __device__ inline size_t head_smem_bytes_arranged(int sig_pad, int kpad_max, int nt) {
    return (size_t)sig_pad * HD_M * 2 + 2 * (size_t)nt * kpad_max * 2 + (size_t)HD_M * (nt * 2 + 16) + 128 + HD_TBL_BYTES + 1024;
}

// Staged rows -> global rows, one warp: 32 rows x up to HC columns.  A lane moves one vector of sizeof(V) / 2 columns; the
// lanes that cover one row's HC columns sit next to each other, so an instruction writes whole contiguous row segments.
// Columns past nvalid are not written (a vector that straddles nvalid falls back to single elements).
template <typename V, int HC>
__device__ __forceinline__ void head_copy_out(const unsigned char* sbase, int spitch, __nv_bfloat16* dbase, int64_t row_stride,
                                              int nrows, int nvalid, int lane) {
    constexpr int VC = sizeof(V) / 2;                       // columns per vector
    constexpr int LPR = HC / VC < 32 ? HC / VC : 32;        // lanes per row
    constexpr int RPI = 32 / LPR;                           // rows per instruction
    constexpr int CPI = LPR * VC;                           // columns per instruction and row
    const int lr = lane / LPR, lc = (lane % LPR) * VC;
#pragma unroll
    for (int c0 = 0; c0 < HC; c0 += CPI) {
        const int c = c0 + lc;
        if (c >= nvalid) continue;
        const bool whole = c + VC <= nvalid;
        const unsigned char* sp = sbase + (size_t)lr * spitch + c * 2;
        __nv_bfloat16* dp = dbase + (size_t)lr * row_stride + c;
        if (whole && nrows == 32) {                              // the common case: all loads in flight, then all stores
            V v[32 / RPI];
#pragma unroll
            for (int i = 0; i < 32 / RPI; ++i) v[i] = *reinterpret_cast<const V*>(sp + (size_t)i * RPI * spitch);
#pragma unroll
            for (int i = 0; i < 32 / RPI; ++i) *reinterpret_cast<V*>(dp + (size_t)i * RPI * row_stride) = v[i];
            continue;
        }
        for (int r = lr; r < nrows; r += RPI, sp += RPI * spitch, dp += RPI * row_stride) {
            if (whole) *reinterpret_cast<V*>(dp) = *reinterpret_cast<const V*>(sp);
            else
                for (int e = 0; e < nvalid - c; ++e) dp[e] = reinterpret_cast<const __nv_bfloat16*>(sp)[e];
        }
    }
}

// A CTA owns 128 positions and a contiguous range of (group, channel-tile) items.  Warp 8 is the producer: it copies
// the item's signal slab into the MN-major A operand (all 32 lanes), streams the packed B tile with cp.async.bulk and
// issues the MMAs (lane 0) into a two-stage TMEM accumulator.  Warps 0-7 drain: quadrant = warp % 4, column half =
// warp / 4; TMEM -> bf16 -> staging rows -> coalesced stores, one item behind the MMAs.
template <int NT, bool ARR>
__global__ void __launch_bounds__(HD_THREADS, 1) signal2weights_tc_kernel(const HeadTCParams p) {
    constexpr int STAGE_PITCH = NT * 2 + 16;
    extern __shared__ unsigned char smem_dyn[];
    // align by pointer arithmetic on the __shared__ array so the compiler keeps the address space (LDS/STS, not generic)
    unsigned char* sm = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#ifdef HSB_HEAD_PROF
    const long long prof_t0 = clock64();
    long long prof_wait = 0, prof_ld = 0, prof_copy = 0, prof_pro = 0, prof_mma_wait = 0;
#endif
    const int kpad = ARR ? p.kpad_max : p.kpad;             // operand stage size; an arranged item may use less
    // arranged mode: one resident A slab (p.kpad = padded width of the head's signal slice) instead of two stages
    const size_t a_bytes = (size_t)(ARR ? p.kpad : kpad) * HD_M * 2, b_bytes = (size_t)NT * kpad * 2;
    unsigned char* a_sm = sm;
    unsigned char* b_sm = a_sm + (ARR ? 1 : 2) * a_bytes;
    unsigned char* st_sm = b_sm + 2 * b_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(st_sm + (size_t)HD_M * STAGE_PITCH);
    uint64_t* b_full = bars;          // [2] B tile landed
    uint64_t* s_empty = bars + 2;     // [2] operand stage (A and B) consumed by the tensor core
    uint64_t* d_full = bars + 4;      // [2] accumulator ready
    uint64_t* d_empty = bars + 6;     // [2] accumulator drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    int4* tbl_sm = reinterpret_cast<int4*>(bars + 16);                          // arranged mode: this CTA's items
    int64_t* pos_sm = reinterpret_cast<int64_t*>(tbl_sm + HD_TBL_ITEMS);        // element offset of each 8-position unit (-1: past the end)

    const int n0 = blockIdx.x * HD_M;
    const int per = (p.items + p.splits - 1) / p.splits;
    const int it0 = blockIdx.y * per, it1 = min(p.items, it0 + per);
    const int nitems = it1 - it0;
    if (nitems <= 0) return;
    auto item = [&](int j) { if (j < HD_TBL_ITEMS) return tbl_sm[j]; int4 e = __ldg(p.table + it0 + j); e.x += p.tbl_base; return e; };

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
    if (ARR) {
        // the head's whole signal slice for these 128 positions, MN-major: unit(mc, k) = (k/8)*LBO + mc*128 + (k%8)*16 holds
        // 8 consecutive positions of channel sig_first + k (sig_first = slice start rounded down to 8); channels outside
        // the slice are zero (their packed weights are zero as well)
        const int sig_first = p.sig_first;
        if (tid < HD_M / 8) {
            const int n = n0 + tid * 8;                                   // P % 8 == 0: a unit never straddles images
            pos_sm[tid] = n < p.NTOT ? (int64_t)(n / p.P) * p.ssb + n % p.P : -1;
        }
        for (int j = tid; j < min(nitems, HD_TBL_ITEMS); j += HD_THREADS) { int4 e = __ldg(p.table + it0 + j); e.x += p.tbl_base; tbl_sm[j] = e; }
        __syncthreads();
        // the packed weights of the first two items do not depend on the slab: their copies fly while it is staged
        if (tid == HD_EPI_WARPS * 32) {
            for (int j = 0; j < min(nitems, 2); ++j) {
                const uint32_t bytes = (uint32_t)((size_t)NT * tbl_sm[j].y * 2);
                mbar_arrive_expect_tx(b_full + j, bytes);
                bulk_g2s(b_sm + j * b_bytes, p.packed + (size_t)(it0 + j) * NT * kpad, bytes, b_full + j);
            }
        }
        // 16-byte chunk L of the slab = (k / 8, mc, k % 8): consecutive threads write consecutive chunks (conflict-free) and
        // read, per channel, 64 contiguous bytes (4 position units)
        // only the 8-channel groups this CTA's items read (a CTA holds a quarter or so of the row's tiles)
        int kg_lo = 0, kg_hi = p.kpad / 8;
        if (nitems <= HD_TBL_ITEMS) {
            int lo = 1 << 30, hi = 0;
            for (int j = 0; j < nitems; ++j) { const int4 e = tbl_sm[j]; lo = min(lo, e.x); hi = max(hi, e.x + e.y); }
            kg_lo = max(0, (lo - sig_first) >> 3); kg_hi = min(kg_hi, (hi - sig_first + 7) >> 3);
        }
        const int chunk0 = kg_lo * HD_M, chunks = kg_hi * HD_M;       // 128 chunks (16 units x 8 channels) per group
        constexpr int DEPTH = 14;                  // 16-byte loads in flight per thread: one round for a 240-channel range
        for (int base = chunk0; base < chunks; base += DEPTH * HD_THREADS) {
            uint4 v[DEPTH];
#pragma unroll
            for (int e = 0; e < DEPTH; ++e) {
                const int L = base + e * HD_THREADS + tid;
                const int k = (L >> 7) * 8 + (L & 7), mc = (L >> 3) & (HD_M / 8 - 1), ch = sig_first + k;
                v[e] = make_uint4(0, 0, 0, 0);
                if (L < chunks && ch >= p.sig_index && ch < p.sig_end) {
                    const int64_t off = pos_sm[mc];
                    if (off >= 0) v[e] = __ldg(reinterpret_cast<const uint4*>(p.s + off + (size_t)ch * p.ssc));
                }
            }
#pragma unroll
            for (int e = 0; e < DEPTH; ++e) {
                const int L = base + e * HD_THREADS + tid;
                if (L < chunks) *reinterpret_cast<uint4*>(a_sm + (size_t)L * 16) = v[e];
            }
        }
        fence_proxy_async_smem();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
#ifdef HSB_HEAD_PROF
    prof_pro = clock64() - prof_t0;
#endif
    const int a_lbo = (HD_M / 8) * 128, b_lbo = (NT / 8) * 128;

    if (warp == HD_EPI_WARPS) {
        // ================= producer / MMA warp =================
        const uint32_t idesc = idesc_bf16_f32(HD_M, NT, /*A MN-major*/ true, false);
        int stage_group0 = -1, stage_group1 = -1;
        auto stage_operands = [&](int j) {          // whole warp: A slab (if the signal range changed) + B tile of item j
            const int st = j & 1, it = it0 + j;
            int g, t, k_first, k_count, k_item;      // slab id, tile, first signal channel, channels to copy, item's padded K
            if (ARR) {
                const int4 e = item(j);
                g = e.x * 4096 + e.y; t = 0; k_first = e.x; k_item = e.y; k_count = max(0, min(e.y, p.sig_end - e.x));
            } else {
                g = it / p.otiles; t = it % p.otiles; k_first = p.sig_index + g * p.spg; k_count = p.spg; k_item = kpad;
            }
            if (j >= 2) mbar_wait(s_empty + st, ((j >> 1) - 1) & 1);      // MMAs of item j-2 done with this stage
            if (!ARR && (st ? stage_group1 : stage_group0) != g) {
                unsigned char* a_dst = a_sm + st * a_bytes;
                // unit(mc, k) = (k/8)*LBO + mc*128 + (k%8)*16: 8 consecutive positions of signal channel k
                // lane = (8-position unit mc = lane % 16, channel parity): position arithmetic once per lane
                const int mc = lane & (HD_M / 8 - 1), n = n0 + mc * 8;
                const bool in_range = n < p.NTOT;
                const int bimg = in_range ? n / p.P : 0, pp = in_range ? n % p.P : 0;   // P % 8 == 0: a unit never straddles images
                const __nv_bfloat16* sbase = p.s + (size_t)bimg * p.ssb + (size_t)k_first * p.ssc + pp;
                for (int kb = lane >> 4; kb < k_item; kb += 16) {       // 8 independent 16-byte loads in flight per lane
                    uint4 v[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int k = kb + 2 * e;
                        v[e] = make_uint4(0, 0, 0, 0);
                        if (k < k_count && in_range) v[e] = __ldg(reinterpret_cast<const uint4*>(sbase + (size_t)k * p.ssc));
                    }
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int k = kb + 2 * e;
                        if (k < k_item) *reinterpret_cast<uint4*>(a_dst + (k >> 3) * a_lbo + mc * 128 + (k & 7) * 16) = v[e];
                    }
                }
                if (st) stage_group1 = g; else stage_group0 = g;
                fence_proxy_async_smem();
            }
            __syncwarp();
            if ((!ARR || j >= 2) && elect_one()) {            // arranged: items 0 and 1 were requested before the slab was staged
                const uint32_t bytes = (uint32_t)((size_t)NT * k_item * 2);
#ifdef HSB_HEAD_NOLOAD      // timing experiment only (wrong results): how much of the item period is the weight-tile copy?
                (void)bytes; mbar_arrive(b_full + st);
#else
                mbar_arrive_expect_tx(b_full + st, bytes);
                bulk_g2s(b_sm + st * b_bytes, p.packed + (ARR ? (size_t)it * NT * kpad : ((size_t)g * p.otiles + t) * NT * kpad), bytes, b_full + st);
#endif
            }
        };
        stage_operands(0);
        for (int j = 0; j < nitems; ++j) {
            const int st = j & 1;
            if (j + 1 < nitems) stage_operands(j + 1);          // overlaps the MMAs / epilogue of item j
            if (elect_one()) {
#ifdef HSB_HEAD_PROF
                const long long tw_ = clock64();
#endif
                mbar_wait(b_full + st, (j >> 1) & 1);
#ifdef HSB_HEAD_PROF
                prof_ld += clock64() - tw_;
                const long long tw2_ = clock64();
#endif
                if (j >= 2) mbar_wait(d_empty + st, ((j >> 1) - 1) & 1);      // epilogue drained this accumulator
#ifdef HSB_HEAD_PROF
                prof_mma_wait += clock64() - tw2_;
#endif
                tc_fence_after_sync();
                // arranged: the item's first channel (a multiple of 8 past the slab's first) selects the k-chunk of the resident slab
                const int4 te = ARR ? item(j) : make_int4(0, kpad, 0, 0);
                const uint32_t a_addr = ARR ? smem_u32(a_sm) + ((te.x - p.sig_first) >> 3) * a_lbo : smem_u32(a_sm + st * a_bytes);
                const uint32_t b_addr = smem_u32(b_sm + st * b_bytes);
                const int ksteps = te.y / 16;
                for (int s = 0; s < ksteps; ++s) {
                    const uint64_t da = smem_desc(a_addr + 2 * s * a_lbo, a_lbo, 128, SWZ_NONE);
                    const uint64_t db = smem_desc(b_addr + 2 * s * b_lbo, b_lbo, 128, SWZ_NONE);
                    umma_bf16(tmem + st * NT, da, db, idesc, s > 0);
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
        constexpr int HC = NT / 2;                        // columns per warp
        unsigned char* my_row = st_sm + (size_t)row * STAGE_PITCH + half * HC * 2;
        for (int j = 0; j < nitems; ++j) {
            const int st = j & 1, it = it0 + j;
            int o_base, o_end;
            if (ARR) {
                const int4 e = item(j);
                o_base = e.z; o_end = min(e.z + e.w, p.hp);
            } else {
                const int g = it / p.otiles, t = it % p.otiles;
                o_base = g * p.opg + t * NT;
                o_end = min(min((g + 1) * p.opg, p.hp), o_base + NT);
            }
            const int nvalid = min(HC, max(o_end - o_base, 0) - half * HC);   // valid columns in this warp's half (may be <= 0)
            const int ncols = min(HC, (max(nvalid, 0) + 31) & ~31);
#ifdef HSB_HEAD_PROF
            const long long tw_ = clock64();
#endif
            mbar_wait(d_full + st, (j >> 1) & 1);
#ifdef HSB_HEAD_PROF
            prof_wait += clock64() - tw_;
            const long long tc_ = clock64();
#endif
            tc_fence_after_sync();
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + st * NT + half * HC;
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
#ifdef HSB_HEAD_PROF
            prof_ld += clock64() - tc_;
            const long long tq_ = clock64();
#endif
            // coalesced write-out of this warp's 32 rows x its column half: V columns per lane, 2 HC / (64 V)... rows per
            // instruction (16-byte lanes: 4 rows of 128 contiguous bytes each); V = the widest vector the start column allows
            if (nvalid > 0) {
                const int ob = o_base + half * HC;
                const int nrows = min(32, p.NTOT - (n0 + q * 32));
                const unsigned char* sbase = st_sm + (size_t)(q * 32) * STAGE_PITCH + half * HC * 2;
                __nv_bfloat16* dbase = p.out + (size_t)(n0 + q * 32) * p.row_stride + ob;
                const bool row16 = (p.row_stride & 7) == 0 && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
                if (row16 && (ob & 7) == 0) head_copy_out<uint4, HC>(sbase, STAGE_PITCH, dbase, p.row_stride, nrows, nvalid, lane);
                else if (row16 && (ob & 3) == 0) head_copy_out<uint2, HC>(sbase, STAGE_PITCH, dbase, p.row_stride, nrows, nvalid, lane);
                else if ((p.row_stride & 1) == 0 && (ob & 1) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 3) == 0)
                    head_copy_out<uint32_t, HC>(sbase, STAGE_PITCH, dbase, p.row_stride, nrows, nvalid, lane);
                else head_copy_out<unsigned short, HC>(sbase, STAGE_PITCH, dbase, p.row_stride, nrows, nvalid, lane);
            }
            __syncwarp();       // staging rows are rewritten by the next item
#ifdef HSB_HEAD_PROF
            prof_copy += clock64() - tq_;
#endif
        }
    }
#ifdef HSB_HEAD_PROF
    if (p.prof && (tid == 0 || tid == HD_EPI_WARPS * 32)) {
        long long* o = p.prof + ((size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 2 + (tid == 0 ? 0 : 1)) * 8;
        o[0] = clock64() - prof_t0; o[1] = prof_pro; o[2] = prof_wait; o[3] = prof_ld; o[4] = prof_copy; o[5] = prof_mma_wait; o[6] = nitems;
    }
#endif
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

// ---- reference-order rows with the signal slice resident (same kernel mode as the arranged rows) -----------------------------
// The K = 12..80 heads spent their time re-staging a 4 KB signal slab per (group, tile) item.  Keeping the head's whole signal
// slice in shared memory (as the arranged heads do) turns an item into "stream one packed weight tile, issue 1-3 MMAs":
// items are 128-column tiles of the reference-order row that may straddle groups (K = union of their signal ranges, zero
// blocks in the packed weights), planned greedily in 8-column units.  The packed buffer carries its own item table:
//   [int4 table[items]] [pad to 128 B] [items x 128 x kpad_max bf16]
constexpr int HD_NT_RO = 128;

struct ResidentPlan {
    int items, kpad_max, sig_pad;
    std::vector<int4> table;               // {first signal channel relative to sig_index (multiple of 8), kpad, first column, columns}
};

static bool plan_resident(int sig_ch, int out_ch, int groups, ResidentPlan* pl) {
    if (sig_ch <= 0 || out_ch <= 0 || groups <= 0 || sig_ch % groups || out_ch % groups) return false;
    const int spg = sig_ch / groups, opg = out_ch / groups;
    pl->sig_pad = (sig_ch + 15) / 16 * 16 + 16;
    pl->kpad_max = 16;
    pl->table.clear();
    const long budget = 227L * 1024 - (long)pl->sig_pad * HD_M * 2 - (long)HD_M * (HD_NT_RO * 2 + 16) - 2048 - HD_TBL_BYTES;
    const int KLIM = (int)std::min<long>(160, budget / (2 * HD_NT_RO * 2) / 16 * 16);     // 160: a tile straddling two 80-channel groups
    if (KLIM < 16) return false;
    auto kfirst_of = [&](int lo) { return (lo * spg) & ~7; };
    auto kpad_of = [&](int lo, int hi) { return ((hi + 1) * spg - kfirst_of(lo) + 15) / 16 * 16; };
    int c0 = 0, gmin = groups, gmax = -1;
    auto flush = [&](int c1) {
        pl->table.push_back(make_int4(kfirst_of(gmin), kpad_of(gmin, gmax), c0, c1 - c0));
        pl->kpad_max = std::max(pl->kpad_max, kpad_of(gmin, gmax));
        c0 = c1; gmin = groups; gmax = -1;
    };
    for (int u = 0; u < out_ch; u += 8) {
        const int ulo = u / opg, uhi = (std::min(u + 8, out_ch) - 1) / opg;
        if (kpad_of(ulo, uhi) > KLIM) return false;              // one unit alone is too wide
        const int nlo = std::min(gmin, ulo), nhi = std::max(gmax, uhi);
        if (u > c0 && (u - c0 >= HD_NT_RO || kpad_of(nlo, nhi) > KLIM)) flush(u);
        gmin = std::min(gmin, ulo); gmax = std::max(gmax, uhi);
    }
    flush(out_ch);
    pl->items = (int)pl->table.size();
    return head_smem_bytes_arranged(pl->sig_pad, pl->kpad_max, HD_NT_RO) <= 227 * 1024;
}

static inline size_t resident_table_bytes(int items) { return ((size_t)items * sizeof(int4) + 127) / 128 * 128; }

// legacy layout (one A slab per item) only where the slice does not fit one CTA; HSB_HEAD_LEGACY=1 forces it (A/B timing)
static bool use_resident(int sig_ch, int out_ch, int groups, ResidentPlan* pl) {
    static const bool legacy = [] { const char* v = getenv("HSB_HEAD_LEGACY"); return v && v[0] == '1'; }();
    return !legacy && plan_resident(sig_ch, out_ch, groups, pl);
}

template <typename T>
__global__ void head_pack_resident_kernel(const T* __restrict__ ws, __nv_bfloat16* __restrict__ out, const int4* __restrict__ table,
                                          const float* __restrict__ scale, int spg, int opg, int out_ch, int items, int kpad_max) {
    const size_t per_item = (size_t)HD_NT_RO * kpad_max;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < per_item * items; i += (size_t)gridDim.x * blockDim.x) {
        const int it = (int)(i / per_item);
        size_t r = i % per_item;
        const int4 e = table[it];
        float v = 0.f;
        if (r < (size_t)HD_NT_RO * e.y) {
            const int k8 = r % 8; r /= 8;
            const int n8 = r % 8; r /= 8;
            const int nc = r % (HD_NT_RO / 8); r /= (HD_NT_RO / 8);
            const int kc = (int)r, col = e.z + nc * 8 + n8;
            if (nc * 8 + n8 < e.w && col < out_ch) {
                const int k = e.x + kc * 8 + k8 - (col / opg) * spg;          // channel inside the group of this output
                if (k >= 0 && k < spg) {
                    v = ld_f(ws + (size_t)col * spg + k);
                    if (scale) v *= scale[col];
                }
            }
        }
        out[i] = __float2bfloat16_rn(v);
    }
}

}  // namespace hsb

using namespace hsb;

extern "C" int64_t hsb_head_packed_elems(int sig_ch, int out_ch, int groups) {
    if (sig_ch <= 0 || out_ch <= 0 || groups <= 0 || sig_ch % groups || out_ch % groups) return -1;
    ResidentPlan pl;
    if (use_resident(sig_ch, out_ch, groups, &pl)) return (int64_t)(resident_table_bytes(pl.items) / 2) + (int64_t)pl.items * HD_NT_RO * pl.kpad_max;
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
    ResidentPlan pl;
    if (use_resident(sig_ch, out_ch, groups, &pl)) {
        HSB_REQUIRE(((uintptr_t)packed % 16) == 0, HSB_ERR_UNSUPPORTED, "head_pack: packed buffer must be 16-byte aligned");
        // one-time operation: the table is tiny, a synchronous copy keeps the host vector's lifetime simple
        cudaError_t ce = cudaMemcpyAsync(packed, pl.table.data(), pl.table.size() * sizeof(int4), cudaMemcpyHostToDevice, st);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
        if (ce != cudaSuccess) return fail(HSB_ERR_CUDA, std::string("head_pack: table copy: ") + cudaGetErrorString(ce));
        __nv_bfloat16* tiles = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<unsigned char*>(packed) + resident_table_bytes(pl.items));
        if (dtype == HSB_F32)
            head_pack_resident_kernel<float><<<blocks, 256, 0, st>>>((const float*)ws, tiles, (const int4*)packed, row_scale, spg, opg, out_ch, pl.items, pl.kpad_max);
        else
            head_pack_resident_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)ws, tiles, (const int4*)packed, row_scale, spg, opg, out_ch,
                                                                             pl.items, pl.kpad_max);
        return check_launch("head_pack (resident) launch");
    }
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
    ResidentPlan pl;
    if (use_resident(sig_ch, out_ch, groups, &pl)) {
        int n_items = 0;                                     // items that start before hp (the kernel clips the last one)
        while (n_items < pl.items && pl.table[n_items].z < hp) ++n_items;
        p.s = (const __nv_bfloat16*)s; p.out = (__nv_bfloat16*)w_out;
        p.table = (const int4*)packed;
        p.packed = reinterpret_cast<const __nv_bfloat16*>(reinterpret_cast<const unsigned char*>(packed) + resident_table_bytes(pl.items));
        p.NTOT = B * P; p.P = P; p.sig_index = sig_index; p.spg = 0; p.kpad = pl.sig_pad; p.opg = 0; p.hp = hp; p.groups = 1; p.otiles = n_items;
        p.ssb = s_stride_b; p.ssc = s_stride_c; p.row_stride = out_row_stride;
        p.kpad_max = pl.kpad_max; p.sig_end = sig_index + sig_ch; p.sig_first = sig_index; p.tbl_base = sig_index; p.prof = nullptr;
        const size_t smem = head_smem_bytes_arranged(pl.sig_pad, pl.kpad_max, HD_NT_RO);
        auto kern = signal2weights_tc_kernel<HD_NT_RO, true>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(HSB_ERR_CUDA, std::string("signal2weights_packed attr: ") + cudaGetErrorString(e));
        p.items = n_items;
        const int tiles = ceil_div(p.NTOT, HD_M), sms = std::max(1, device_sm_count());
        int splits = 1, best = ceil_div(tiles, sms) * (p.items + 4);
        for (int sp = 2; sp <= p.items; ++sp) {
            const int cost = ceil_div(tiles * sp, sms) * (ceil_div(p.items, sp) + 4);
            if (cost < best) { best = cost; splits = sp; }
        }
        p.splits = splits;
        HSB_REQUIRE(splits <= 65535, HSB_ERR_UNSUPPORTED, "signal2weights_packed: grid too large");
        kern<<<dim3(tiles, splits), HD_THREADS, smem, (cudaStream_t)stream>>>(p);
        note_kernel("signal2weights_tc_kernel<resident>");
        return check_launch("signal2weights_packed launch");
    }
    p.s = (const __nv_bfloat16*)s; p.packed = (const __nv_bfloat16*)packed; p.out = (__nv_bfloat16*)w_out;
    p.NTOT = B * P; p.P = P; p.sig_index = sig_index; p.spg = sig_ch / groups; p.kpad = (p.spg + 15) / 16 * 16;
    p.opg = out_ch / groups; p.hp = hp; p.groups = groups; p.otiles = ceil_div(p.opg, HD_NT);
    p.ssb = s_stride_b; p.ssc = s_stride_c; p.row_stride = out_row_stride;
    const size_t smem = head_smem_bytes(p.kpad);
    HSB_REQUIRE(smem <= 227 * 1024, HSB_ERR_UNSUPPORTED, "signal2weights_packed: sig_ch / groups too large for one CTA");
    cudaError_t e = cudaFuncSetAttribute(signal2weights_tc_kernel<HD_NT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
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
    p.table = nullptr; p.kpad_max = p.kpad; p.sig_end = 0; p.prof = nullptr; p.sig_first = 0; p.tbl_base = 0;
    signal2weights_tc_kernel<HD_NT, false><<<grid, HD_THREADS, smem, (cudaStream_t)stream>>>(p);
    note_kernel("signal2weights_tc_kernel");
    return check_launch("signal2weights_packed launch");
}

// ---- arranged rows for the fused inverted-residual MetaBlock -------------------------------------------------------------
namespace hsb {

constexpr int HD_NT_ARR = 128;             // arranged mode: narrower tiles keep the signal range of a tile to <= 2-3 groups

struct ArrangedPlan {
    int row_elems, items, kpad_max, sig_pad;
    std::vector<int4> table;               // {first signal channel (relative to the head's slice), kpad, first column, columns}
};

static bool plan_arranged(int sig_index, int sig_ch, int out_ch, int groups, int hp_offset, int cin, int hid, int cout, ArrangedPlan* pl) {
    if (sig_ch <= 0 || out_ch <= 0 || groups <= 0 || sig_ch % groups || out_ch % groups || hp_offset < 0) return false;
    const int hp = cin * hid + 9 * hid + hid * cout;
    if (cin <= 0 || hid <= 0 || cout <= 0 || hp_offset + hp > out_ch) return false;
    const int spg = sig_ch / groups, opg = out_ch / groups;
    pl->row_elems = IRRow(cin, hid, cout).bytes / 2;
    pl->kpad_max = 16;
    pl->table.clear();
    // Greedy tiling of the arranged row in units of 8 columns (one 16-byte operand row: consecutive head outputs): a
    // tile grows to 128 columns unless the next unit would widen its range of signal channels beyond what two operand
    // stages of a CTA can hold -- that happens where the arranged order wraps around (end of one 8-channel chunk of B1,
    // start of the next), because the head's groups are contiguous in the reference order, not in this one.
    // the signal slice stays resident (padded: first channel rounded down to 8, + 16 so that a tile's padded range never
    // leaves the slab); what is left of the shared memory holds two stages of packed weights
    pl->sig_pad = ((sig_index & 7) + sig_ch + 15) / 16 * 16 + 16;
    const long budget = 227L * 1024 - (long)pl->sig_pad * HD_M * 2 - (long)HD_M * (HD_NT_ARR * 2 + 16) - 2048 - HD_TBL_BYTES;
    const int KLIM = (int)std::min<long>(256, budget / (2 * HD_NT_ARR * 2) / 16 * 16);
    if (KLIM < 16) return false;
    int c0 = 0, gmin = groups, gmax = -1;
    // channels relative to the slice; the range starts at a multiple of 8 of the absolute channel index
    auto kfirst_of = [&](int lo) { return ((sig_index + lo * spg) & ~7) - sig_index; };
    auto kpad_of = [&](int lo, int hi) { return hi < 0 ? 16 : ((hi + 1) * spg - kfirst_of(lo) + 15) / 16 * 16; };
    auto flush = [&](int c1) {
        pl->table.push_back(make_int4(gmax < 0 ? kfirst_of(0) : kfirst_of(gmin), kpad_of(gmin, gmax), c0, c1 - c0));
        pl->kpad_max = std::max(pl->kpad_max, kpad_of(gmin, gmax));
        c0 = c1; gmin = groups; gmax = -1;
    };
    for (int u = 0; u < pl->row_elems; u += 8) {
        int ulo = groups, uhi = -1;
        for (int e = u; e < std::min(u + 8, pl->row_elems); ++e) {
            const IRSource src = ir_arranged_source(e, cin, hid, cout);
            if (src.src < 0) continue;
            const int g = (hp_offset + src.src) / opg;
            ulo = std::min(ulo, g); uhi = std::max(uhi, g);
        }
        if (kpad_of(ulo, uhi) > KLIM) return false;              // one unit alone is too wide
        const int nlo = std::min(gmin, ulo), nhi = std::max(gmax, uhi);
        if (u > c0 && (u - c0 >= HD_NT_ARR || kpad_of(nlo, nhi) > KLIM)) flush(u);
        gmin = std::min(gmin, ulo); gmax = std::max(gmax, uhi);
    }
    flush(pl->row_elems);
    pl->items = (int)pl->table.size();
    return true;
}

template <typename T>
__global__ void head_pack_arranged_kernel(const T* __restrict__ ws, __nv_bfloat16* __restrict__ out, const int4* __restrict__ table,
                                          const float* s1, const float* s2, const float* s3, int sig_index, int spg, int opg,
                                          int hp_offset, int cin, int hid, int cout, int items, int kpad_max, int row_elems) {
    const size_t per_item = (size_t)HD_NT_ARR * kpad_max;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < per_item * items; i += (size_t)gridDim.x * blockDim.x) {
        const int it = (int)(i / per_item);
        size_t r = i % per_item;
        const int4 e = table[it];                     // kstart here is absolute (sig_index already added)
        float v = 0.f;
        if (r < (size_t)HD_NT_ARR * e.y) {
            const int k8 = r % 8; r /= 8;
            const int n8 = r % 8; r /= 8;
            const int nc = r % (HD_NT_ARR / 8); r /= (HD_NT_ARR / 8);
            const int kc = (int)r;
            const int col = e.z + nc * 8 + n8;
            if (nc * 8 + n8 < e.w && col < row_elems) {
                const IRSource src = ir_arranged_source(col, cin, hid, cout);
                if (src.src >= 0) {
                    const int o = hp_offset + src.src, g = o / opg;
                    const int k = (e.x - sig_index) + kc * 8 + k8 - g * spg;      // channel inside the group of this output (e.x may start before the slice)
                    if (k >= 0 && k < spg) {
                        const float sc = src.which == 0 ? s1[src.ch] : (src.which == 1 ? s2[src.ch] : s3[src.ch]);
                        v = ld_f(ws + (size_t)o * spg + k) * sc;
                    }
                }
            }
        }
        out[i] = __float2bfloat16_rn(v);
    }
}

}  // namespace hsb

extern "C" int hsb_head_arranged_plan(int sig_index, int sig_ch, int out_ch, int groups, int hp_offset, int Cin, int hid, int Cout,
                                      int64_t* packed_elems, int* n_items, int* kpad_max) {
    ArrangedPlan pl;
    HSB_REQUIRE(plan_arranged(sig_index, sig_ch, out_ch, groups, hp_offset, Cin, hid, Cout, &pl), HSB_ERR_UNSUPPORTED,
                "head_arranged_plan: bad dimensions, or the head's signal slice / a tile's signal range does not fit one CTA");
    HSB_REQUIRE(head_smem_bytes_arranged(pl.sig_pad, pl.kpad_max, HD_NT_ARR) <= 227 * 1024, HSB_ERR_UNSUPPORTED,
                "head_arranged_plan: the head's signal slice is too wide to stay resident in one CTA");
    if (packed_elems) *packed_elems = (int64_t)pl.items * HD_NT_ARR * pl.kpad_max;
    if (n_items) *n_items = pl.items;
    if (kpad_max) *kpad_max = pl.kpad_max;
    return HSB_OK;
}

extern "C" int hsb_head_pack_arranged(const void* ws, void* packed, void* table, const float* bn1_scale, const float* bn2_scale,
                                      const float* bn3_scale, int sig_index, int sig_ch, int out_ch, int groups, int hp_offset,
                                      int Cin, int hid, int Cout, int dtype, void* stream) {
    HSB_REQUIRE(ws && packed && table && bn1_scale && bn2_scale && bn3_scale, HSB_ERR_INVALID_ARG, "head_pack_arranged: null pointer");
    HSB_REQUIRE(dtype == HSB_F32 || dtype == HSB_BF16, HSB_ERR_INVALID_ARG, "head_pack_arranged: bad dtype");
    ArrangedPlan pl;
    HSB_REQUIRE(plan_arranged(sig_index, sig_ch, out_ch, groups, hp_offset, Cin, hid, Cout, &pl), HSB_ERR_UNSUPPORTED, "head_pack_arranged: bad dimensions");
    for (auto& e : pl.table) e.x += sig_index;         // the kernels want absolute signal channels
    cudaStream_t st = (cudaStream_t)stream;
    // one-time operation: the table is tiny, a synchronous copy keeps the host vector's lifetime simple
    cudaError_t ce = cudaMemcpyAsync(table, pl.table.data(), pl.table.size() * sizeof(int4), cudaMemcpyHostToDevice, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    if (ce != cudaSuccess) return fail(HSB_ERR_CUDA, std::string("head_pack_arranged: table copy: ") + cudaGetErrorString(ce));
    const int spg = sig_ch / groups, opg = out_ch / groups;
    const int blocks = std::max(1, device_sm_count()) * 4;
    if (dtype == HSB_F32)
        head_pack_arranged_kernel<float><<<blocks, 256, 0, st>>>((const float*)ws, (__nv_bfloat16*)packed, (const int4*)table, bn1_scale, bn2_scale,
                                                                 bn3_scale, sig_index, spg, opg, hp_offset, Cin, hid, Cout, pl.items, pl.kpad_max, pl.row_elems);
    else
        head_pack_arranged_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)ws, (__nv_bfloat16*)packed, (const int4*)table, bn1_scale,
                                                                         bn2_scale, bn3_scale, sig_index, spg, opg, hp_offset, Cin, hid, Cout, pl.items,
                                                                         pl.kpad_max, pl.row_elems);
    return check_launch("head_pack_arranged launch");
}

extern "C" int hsb_signal2weights_arranged_fwd(const void* s, const void* packed, const void* table, void* w_arranged,
                                               int B, int sig_index, int sig_ch, int n_items, int kpad_max, int row_elems,
                                               int fh, int fw, int64_t s_stride_b, int64_t s_stride_c, int64_t out_row_stride,
                                               void* stream) {
    HSB_REQUIRE(s && packed && table && w_arranged, HSB_ERR_INVALID_ARG, "signal2weights_arranged: null pointer");
    HSB_REQUIRE(B > 0 && sig_ch > 0 && n_items > 0 && kpad_max > 0 && kpad_max % 16 == 0 && row_elems > 0 && fh > 0 && fw > 0 && sig_index >= 0,
                HSB_ERR_INVALID_ARG, "signal2weights_arranged: bad dimension");
    const int P = fh * fw;
    HSB_REQUIRE(P % 8 == 0, HSB_ERR_UNSUPPORTED, "signal2weights_arranged: fh*fw must be a multiple of 8");
    HSB_REQUIRE(out_row_stride >= row_elems, HSB_ERR_INVALID_ARG, "signal2weights_arranged: row stride shorter than the arranged row");
    HSB_REQUIRE(((uintptr_t)s % 16) == 0 && (s_stride_b % 8) == 0 && (s_stride_c % 8) == 0 && ((uintptr_t)packed % 16) == 0 &&
                ((uintptr_t)table % 16) == 0, HSB_ERR_UNSUPPORTED, "signal2weights_arranged: signal / packed weights must be 16-byte aligned");
    HeadTCParams p;
    p.s = (const __nv_bfloat16*)s; p.packed = (const __nv_bfloat16*)packed; p.out = (__nv_bfloat16*)w_arranged;
    const int sig_pad = ((sig_index & 7) + sig_ch + 15) / 16 * 16 + 16;
    p.NTOT = B * P; p.P = P; p.sig_index = sig_index; p.spg = 0; p.kpad = sig_pad; p.opg = 0; p.hp = row_elems; p.groups = 1; p.otiles = n_items;
    p.ssb = s_stride_b; p.ssc = s_stride_c; p.row_stride = out_row_stride;
    p.table = (const int4*)table; p.kpad_max = kpad_max; p.sig_end = sig_index + sig_ch; p.sig_first = sig_index & ~7; p.tbl_base = 0;
    const size_t smem = head_smem_bytes_arranged(sig_pad, kpad_max, HD_NT_ARR);
    HSB_REQUIRE(smem <= 227 * 1024, HSB_ERR_UNSUPPORTED, "signal2weights_arranged: the signal slice does not fit one CTA");
    auto kern = signal2weights_tc_kernel<HD_NT_ARR, true>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(HSB_ERR_CUDA, std::string("signal2weights_arranged attr: ") + cudaGetErrorString(e));
    p.items = n_items;
    const int tiles = ceil_div(p.NTOT, HD_M);
    const int sms = std::max(1, device_sm_count());
    int splits = 1, best = ceil_div(tiles, sms) * (p.items + 4);
    for (int sp = 2; sp <= p.items; ++sp) {
        const int cost = ceil_div(tiles * sp, sms) * (ceil_div(p.items, sp) + 4);
        if (cost < best) { best = cost; splits = sp; }
    }
    p.splits = splits;
    dim3 grid(tiles, splits);
    p.prof = nullptr;
#ifdef HSB_HEAD_PROF
    {
        static long long* dprof = nullptr;
        const int ctas = tiles * splits;
        if (!dprof) cudaMalloc(&dprof, 4096 * 16 * sizeof(long long));
        cudaMemsetAsync(dprof, 0, 4096 * 16 * sizeof(long long), (cudaStream_t)stream);
        p.prof = dprof;
        kern<<<grid, HD_THREADS, smem, (cudaStream_t)stream>>>(p);
        cudaStreamSynchronize((cudaStream_t)stream);
        static long long host[4096 * 16];
        cudaMemcpy(host, dprof, sizeof(long long) * ctas * 16, cudaMemcpyDeviceToHost);
        double a[2][8] = {{0}};
        for (int c = 0; c < ctas; ++c) for (int r = 0; r < 2; ++r) for (int k = 0; k < 8; ++k) a[r][k] += (double)host[(c * 2 + r) * 8 + k] / ctas;
        fprintf(stderr, "[hsb-prof] head<arranged> grid %d x %d, %.1f items per CTA, smem %zu\n", tiles, splits, a[0][6], smem);
        fprintf(stderr, "[hsb-prof]  epilogue warp 0: total %.0f cycles, prologue %.0f, wait d_full %.0f, TMEM->staging %.0f, copy out %.0f\n", a[0][0], a[0][1], a[0][2], a[0][3], a[0][4]);
        fprintf(stderr, "[hsb-prof]  producer lane : total %.0f cycles, prologue %.0f, wait b_full %.0f, wait d_empty %.0f\n", a[1][0], a[1][1], a[1][3], a[1][5]);
        note_kernel("signal2weights_tc_kernel<arranged>");
        return check_launch("signal2weights_arranged launch");
    }
#endif
    kern<<<grid, HD_THREADS, smem, (cudaStream_t)stream>>>(p);
    note_kernel("signal2weights_tc_kernel<arranged>");
    return check_launch("signal2weights_arranged launch");
}
