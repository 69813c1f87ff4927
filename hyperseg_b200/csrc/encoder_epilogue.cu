// Channels-last epilogues of the static encoder convolutions (engine path; outside the decoder hot path).
//
// With eval-mode BatchNorms folded into the preceding convolution, every MBConv stage of the stock encoder
// (reference hyperseg/models/backbones/efficientnet.py:82-123, :275, :289) is conv -> +shift -> swish [-> mean over
// H,W -> SE gate -> x * sigmoid(gate)] [-> + skip].  Stock PyTorch runs the shift, the swish, the mean, the sigmoid, the
// gate product and the skip add as six separate elementwise / reduce kernels over the activation; here they are two
// passes, 16 bytes per thread, over NHWC rows:
//   bias_act_nhwc     y = act(x + bias[c]) (+ residual), in place if wanted, and (optionally) per-CTA partial sums of
//                     the rounded output for the squeeze-and-excitation mean (fixed summation order: deterministic)
//   channel_gate_nhwc y = x * sigmoid(gate[n, c])
#include "common.cuh"

namespace hsb {

constexpr int ROWS = 2;       // rows (16-byte loads) in flight per thread

template <typename T> struct Vec16;
template <> struct Vec16<float> { static constexpr int N = 4; };
template <> struct Vec16<__nv_bfloat16> { static constexpr int N = 8; };

template <typename T, int N>
__device__ __forceinline__ void unpack16(const uint4& raw, float (&v)[N]) {
    if constexpr (N == 4) {
        v[0] = __uint_as_float(raw.x); v[1] = __uint_as_float(raw.y); v[2] = __uint_as_float(raw.z); v[3] = __uint_as_float(raw.w);
    } else {
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) { v[2 * e] = __uint_as_float(w[e] << 16); v[2 * e + 1] = __uint_as_float(w[e] & 0xFFFF0000u); }
    }
}

// rounds to T; `v` is overwritten with the rounded values (what a later reader of y sees)
template <typename T, int N>
__device__ __forceinline__ uint4 pack16(float (&v)[N]) {
    if constexpr (N == 4) {
        return make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
    } else {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
            w[e] = *reinterpret_cast<uint32_t*>(&t);
            v[2 * e] = __uint_as_float(w[e] << 16); v[2 * e + 1] = __uint_as_float(w[e] & 0xFFFF0000u);
        }
        return make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// The pass is instruction-issue bound before it is HBM bound (ncu: 31 instructions per element with IEEE exp +
// division), so swish uses the SFU approximations: ex2.approx + rcp.approx, relative error ~2e-7 -- far inside bf16.
__device__ __forceinline__ float fast_sigmoid(float v) { return __frcp_rn(1.f + __expf(-v)); }

__device__ __forceinline__ float epilogue_act(float v, int act) {
    if (act == HSB_ACT_SILU) return v * fast_sigmoid(v);
    return act_apply(v, act);
}

struct BiasActParams {
    const void* x; const float* bias; const void* res; void* y; float* pool;
    int HW, C, G, L, rows_per_cta, chunks, act;
};

// ACT: a compile-time hsb_act, or -1 for "read p.act" (the activations the encoder does not use); RES / POOL switch
// the skip add and the partial sums: the inner loop carries no run-time mode tests
template <typename T, int ACT, bool RES, bool POOL>
__global__ void bias_act_nhwc_kernel(const BiasActParams p) {
    constexpr int N = Vec16<T>::N;
    extern __shared__ float red[];                       // [L][C] partial sums
    const int tid = threadIdx.x, g = tid % p.G, l = tid / p.G;
    const int n = blockIdx.y, chunk = blockIdx.x;
    const size_t base = (size_t)n * p.HW * p.C + (size_t)g * N;
    const T* x = reinterpret_cast<const T*>(p.x) + base;
    const T* res = RES ? reinterpret_cast<const T*>(p.res) + base : nullptr;
    T* y = reinterpret_cast<T*>(p.y) + base;
    float b[N], acc[N];
#pragma unroll
    for (int e = 0; e < N; ++e) { b[e] = p.bias ? p.bias[g * N + e] : 0.f; acc[e] = 0.f; }
    const int r_end = min(p.HW, (chunk + 1) * p.rows_per_cta);
    // ROWS independent 16-byte loads in flight per thread: the pass is HBM-bound and latency decides the bandwidth
    for (int r0 = chunk * p.rows_per_cta + l; r0 < r_end; r0 += ROWS * p.L) {
        uint4 raw[ROWS], rraw[ROWS];
#pragma unroll
        for (int u = 0; u < ROWS; ++u) {
            const int r = r0 + u * p.L;
            if (r < r_end) {
                raw[u] = *reinterpret_cast<const uint4*>(x + (size_t)r * p.C);
                if (RES) rraw[u] = *reinterpret_cast<const uint4*>(res + (size_t)r * p.C);
            }
        }
#pragma unroll
        for (int u = 0; u < ROWS; ++u) {
            const int r = r0 + u * p.L;
            if (r < r_end) {
                float v[N];
                unpack16<T, N>(raw[u], v);
#pragma unroll
                for (int e = 0; e < N; ++e) v[e] = epilogue_act(v[e] + b[e], ACT >= 0 ? ACT : p.act);
                if (RES) {
                    float s[N];
                    unpack16<T, N>(rraw[u], s);
#pragma unroll
                    for (int e = 0; e < N; ++e) v[e] += s[e];
                }
                *reinterpret_cast<uint4*>(y + (size_t)r * p.C) = pack16<T, N>(v);
                if (POOL) {
#pragma unroll
                    for (int e = 0; e < N; ++e) acc[e] += v[e];
                }
            }
        }
    }
    if (POOL) {
#pragma unroll
        for (int e = 0; e < N; ++e) red[l * p.C + g * N + e] = acc[e];
        __syncthreads();
        for (int c = tid; c < p.C; c += blockDim.x) {
            float s = 0.f;
            for (int i = 0; i < p.L; ++i) s += red[i * p.C + c];
            p.pool[((size_t)n * p.chunks + chunk) * p.C + c] = s;
        }
    }
}

struct GateParams { const void* x; const void* gate; void* y; int HW, C, G, L, rows_per_cta; };

template <typename T>
__global__ void channel_gate_nhwc_kernel(const GateParams p) {
    constexpr int N = Vec16<T>::N;
    const int tid = threadIdx.x, g = tid % p.G, l = tid / p.G;
    const int n = blockIdx.y, chunk = blockIdx.x;
    const size_t base = (size_t)n * p.HW * p.C + (size_t)g * N;
    const T* x = reinterpret_cast<const T*>(p.x) + base;
    T* y = reinterpret_cast<T*>(p.y) + base;
    float s[N];
    unpack16<T, N>(*reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(p.gate) + (size_t)n * p.C + (size_t)g * N), s);
#pragma unroll
    for (int e = 0; e < N; ++e) s[e] = fast_sigmoid(s[e]);
    if constexpr (N == 8) {                               // torch rounds sigmoid(gate) to bf16 before the product
        (void)pack16<T, N>(s);
    }
    const int r_end = min(p.HW, (chunk + 1) * p.rows_per_cta);
    for (int r0 = chunk * p.rows_per_cta + l; r0 < r_end; r0 += ROWS * p.L) {
        uint4 raw[ROWS];
#pragma unroll
        for (int u = 0; u < ROWS; ++u)
            if (r0 + u * p.L < r_end) raw[u] = *reinterpret_cast<const uint4*>(x + (size_t)(r0 + u * p.L) * p.C);
#pragma unroll
        for (int u = 0; u < ROWS; ++u) {
            if (r0 + u * p.L < r_end) {
                float v[N];
                unpack16<T, N>(raw[u], v);
#pragma unroll
                for (int e = 0; e < N; ++e) v[e] *= s[e];
                *reinterpret_cast<uint4*>(y + (size_t)(r0 + u * p.L) * p.C) = pack16<T, N>(v);
            }
        }
    }
}

constexpr int MAX_CHUNKS = 64;

// thread arrangement shared by both kernels: G = C/N channel groups x L row lanes, <= 256 threads
static void arrange(int C, int vec, int HW, int& G, int& L, int& rows_per_cta, int& chunks) {
    G = C / vec;
    L = G >= 256 ? 1 : 256 / G;
    if (L > HW) L = HW;
    rows_per_cta = L * 16;
    const int cap = (HW + MAX_CHUNKS - 1) / MAX_CHUNKS;          // at most MAX_CHUNKS partial sums per image
    if (rows_per_cta < cap) rows_per_cta = (cap + L - 1) / L * L;
    chunks = (HW + rows_per_cta - 1) / rows_per_cta;
}

}  // namespace hsb

using namespace hsb;

extern "C" int hsb_bias_act_nhwc_chunks(int C, int64_t HW, int dtype) {
    const int vec = dtype == HSB_BF16 ? 8 : 4;
    if (C <= 0 || HW <= 0 || C % vec != 0 || C / vec > 1024) return -1;
    int G, L, rows, chunks;
    arrange(C, vec, (int)HW, G, L, rows, chunks);
    return chunks;
}

extern "C" int hsb_bias_act_nhwc_fwd(const void* x, const float* bias, const void* residual, void* y, float* pool_partial,
                                     int N, int64_t HW, int C, int act, int dtype, void* stream) {
    HSB_REQUIRE(x && y, HSB_ERR_INVALID_ARG, "bias_act_nhwc: null tensor");
    HSB_REQUIRE(dtype == HSB_F32 || dtype == HSB_BF16, HSB_ERR_INVALID_ARG, "bias_act_nhwc: dtype");
    HSB_REQUIRE(act >= HSB_ACT_NONE && act <= HSB_ACT_SILU, HSB_ERR_INVALID_ARG, "bias_act_nhwc: act");
    const int vec = dtype == HSB_BF16 ? 8 : 4;
    HSB_REQUIRE(N > 0 && HW > 0 && HW < (1 << 30) && C > 0 && C % vec == 0 && C / vec <= 1024, HSB_ERR_UNSUPPORTED,
                "bias_act_nhwc: C must be a multiple of 16 bytes (and at most 1024 vectors)");
    HSB_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(residual)) & 15) == 0,
                HSB_ERR_UNSUPPORTED, "bias_act_nhwc: tensors must be 16-byte aligned");
    BiasActParams p;
    p.x = x; p.bias = bias; p.res = residual; p.y = y; p.pool = pool_partial; p.HW = (int)HW; p.C = C; p.act = act;
    arrange(C, vec, p.HW, p.G, p.L, p.rows_per_cta, p.chunks);
    const dim3 grid(p.chunks, N), block(p.G * p.L);
    const size_t smem = pool_partial ? (size_t)p.L * C * sizeof(float) : 0;
    HSB_REQUIRE(smem <= 48 * 1024, HSB_ERR_UNSUPPORTED, "bias_act_nhwc: channel count too large for the pooled variant");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool bf = dtype == HSB_BF16, has_res = residual != nullptr, has_pool = pool_partial != nullptr;
#define HSB_EPI_LAUNCH(T, A)                                                                              \
    do {                                                                                                  \
        if (has_res && has_pool) bias_act_nhwc_kernel<T, A, true, true><<<grid, block, smem, st>>>(p);    \
        else if (has_res) bias_act_nhwc_kernel<T, A, true, false><<<grid, block, smem, st>>>(p);          \
        else if (has_pool) bias_act_nhwc_kernel<T, A, false, true><<<grid, block, smem, st>>>(p);         \
        else bias_act_nhwc_kernel<T, A, false, false><<<grid, block, smem, st>>>(p);                      \
    } while (0)
    if (act == HSB_ACT_SILU) { if (bf) HSB_EPI_LAUNCH(__nv_bfloat16, HSB_ACT_SILU); else HSB_EPI_LAUNCH(float, HSB_ACT_SILU); }
    else if (act == HSB_ACT_NONE) { if (bf) HSB_EPI_LAUNCH(__nv_bfloat16, HSB_ACT_NONE); else HSB_EPI_LAUNCH(float, HSB_ACT_NONE); }
    else { if (bf) HSB_EPI_LAUNCH(__nv_bfloat16, -1); else HSB_EPI_LAUNCH(float, -1); }
#undef HSB_EPI_LAUNCH
    return check_launch("bias_act_nhwc launch");
}

extern "C" int hsb_channel_gate_nhwc_fwd(const void* x, const void* gate, void* y, int N, int64_t HW, int C, int dtype,
                                         void* stream) {
    HSB_REQUIRE(x && y && gate, HSB_ERR_INVALID_ARG, "channel_gate_nhwc: null tensor");
    HSB_REQUIRE(dtype == HSB_F32 || dtype == HSB_BF16, HSB_ERR_INVALID_ARG, "channel_gate_nhwc: dtype");
    const int vec = dtype == HSB_BF16 ? 8 : 4;
    HSB_REQUIRE(N > 0 && HW > 0 && HW < (1 << 30) && C > 0 && C % vec == 0 && C / vec <= 1024, HSB_ERR_UNSUPPORTED,
                "channel_gate_nhwc: C must be a multiple of 16 bytes (and at most 1024 vectors)");
    HSB_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(gate)) & 15) == 0,
                HSB_ERR_UNSUPPORTED, "channel_gate_nhwc: tensors must be 16-byte aligned");
    GateParams p;
    p.x = x; p.gate = gate; p.y = y; p.HW = (int)HW; p.C = C;
    int chunks;
    arrange(C, vec, p.HW, p.G, p.L, p.rows_per_cta, chunks);
    const dim3 grid(chunks, N), block(p.G * p.L);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == HSB_BF16) channel_gate_nhwc_kernel<__nv_bfloat16><<<grid, block, 0, st>>>(p);
    else channel_gate_nhwc_kernel<float><<<grid, block, 0, st>>>(p);
    return check_launch("channel_gate_nhwc launch");
}
