// Weight head: grouped 1x1 convolution signal -> per-patch weights, written patch-major.
//
// Replaces apply_signal2weights + the grouped nn.Conv2d (reference hyperseg/models/hyperseg_v1_0.py:315-326,
// :473-484), WeightLayer.forward (hyperseg_v1_0_unify.py:287-309) and each head of Conv2dMulti
// (hyperseg_v0_1.py:336-362).  Per group this is a small-K GEMM
//     Wout[n, o] = sum_k Ws[o, k] * s[n, idx + g*K + k],   n = (b, i, j) position, K = sig_ch / groups
// whose cost is writing Wout (15.9 MB per 512x1024 image, bf16); the output goes straight into the
// (B, fh, fw, row) layout the patch kernels stream, which removes the reference's
// permute(0,2,3,1).reshape copy (hyperseg_v1_0.py:345-347).
//
// CTA tile: 32 positions x 128 output channels of one group; thread tile 4 positions x 4 channels
// (lane -> channel quad so a warp writes 256 contiguous bytes of one weight row).
#include "common.cuh"

namespace hsb {

constexpr int S2W_NP = 32;    // positions per CTA
constexpr int S2W_OT = 128;   // output channels per CTA
constexpr int S2W_KC = 32;    // K chunk staged per iteration
constexpr int S2W_OTP = S2W_OT + 4;

struct S2WParams {
    const void* s; const void* ws; void* out;
    int B, P, sig_index, sig_ch, out_ch, hp, groups, spg, opg, otiles;
    int64_t ssb, ssc, ssp;        // signal strides
    int64_t osb, osp, osk;        // output strides (image, patch, weight index)
};

template <typename T>
__global__ void __launch_bounds__(256) signal2weights_kernel(const S2WParams p) {
    __shared__ __align__(16) float s_sm[S2W_KC][S2W_NP];
    __shared__ __align__(16) float w_sm[S2W_KC][S2W_OTP];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = blockIdx.y / p.otiles, ot = blockIdx.y % p.otiles;
    const int n0 = blockIdx.x * S2W_NP;
    const int NT = p.B * p.P;
    const int o_base = g * p.opg + ot * S2W_OT;                 // first output channel of this tile
    const int o_end = min(min((g + 1) * p.opg, p.hp), o_base + S2W_OT);
    const int n_out = o_end - o_base;                           // may be <= 0 for fully padded tiles
    if (n_out <= 0) return;
    const T* s = reinterpret_cast<const T*>(p.s);
    const T* ws = reinterpret_cast<const T*>(p.ws);
    T* out = reinterpret_cast<T*>(p.out);

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < p.spg; k0 += S2W_KC) {
        const int kc = min(S2W_KC, p.spg - k0);
        // signal chunk [kc][NP]
        for (int idx = tid; idx < kc * S2W_NP; idx += blockDim.x) {
            int k = idx / S2W_NP, q = idx % S2W_NP;
            int n = n0 + q;
            float v = 0.f;
            if (n < NT) {
                int b = n / p.P, pp = n % p.P;
                v = ld_f(s + (size_t)b * p.ssb + (size_t)(p.sig_index + g * p.spg + k0 + k) * p.ssc +
                         (size_t)pp * p.ssp);
            }
            s_sm[k][q] = v;
        }
        // static weights chunk, transposed to [kc][OT]
        for (int idx = tid; idx < n_out * kc; idx += blockDim.x) {
            int ol = idx / kc, k = idx % kc;
            w_sm[k][ol] = ld_f(ws + (size_t)(o_base + ol) * p.spg + k0 + k);
        }
        __syncthreads();
        for (int k = 0; k < kc; ++k) {
            const float4 wv = *reinterpret_cast<const float4*>(&w_sm[k][lane * 4]);
            const float4 sv = *reinterpret_cast<const float4*>(&s_sm[k][warp * 4]);
            const float wa[4] = {wv.x, wv.y, wv.z, wv.w};
            const float sa[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(sa[i], wa[j], acc[i][j]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = n0 + warp * 4 + i;
        if (n >= NT) continue;
        const int b = n / p.P, pp = n % p.P;
        T* row = out + (size_t)b * p.osb + (size_t)pp * p.osp;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int ol = lane * 4 + j;
            if (ol < n_out) st_f(row + (size_t)(o_base + ol) * p.osk, acc[i][j]);
        }
    }
}

}  // namespace hsb

using namespace hsb;

extern "C" int hsb_signal2weights_fwd(const void* s, const void* ws, void* w_out,
                                      int B, int sig_index, int sig_ch, int out_ch, int hp, int groups,
                                      int fh, int fw,
                                      int64_t s_stride_b, int64_t s_stride_c, int64_t s_stride_p,
                                      int dtype, int out_layout, int64_t out_row_stride, void* stream) {
    HSB_REQUIRE(s && ws && w_out, HSB_ERR_INVALID_ARG, "signal2weights: null pointer");
    HSB_REQUIRE(B > 0 && sig_ch > 0 && out_ch > 0 && hp > 0 && groups > 0 && fh > 0 && fw > 0 && sig_index >= 0,
                HSB_ERR_INVALID_ARG, "signal2weights: bad dimension");
    HSB_REQUIRE(sig_ch % groups == 0 && out_ch % groups == 0, HSB_ERR_INVALID_ARG,
                "signal2weights: channels not divisible by groups");
    HSB_REQUIRE(hp <= out_ch, HSB_ERR_INVALID_ARG, "signal2weights: hp > out_ch");
    HSB_REQUIRE(dtype == HSB_F32 || dtype == HSB_BF16, HSB_ERR_INVALID_ARG, "signal2weights: bad dtype");
    S2WParams p;
    p.s = s; p.ws = ws; p.out = w_out;
    p.B = B; p.P = fh * fw; p.sig_index = sig_index; p.sig_ch = sig_ch; p.out_ch = out_ch; p.hp = hp;
    p.groups = groups; p.spg = sig_ch / groups; p.opg = out_ch / groups;
    p.otiles = ceil_div(p.opg, S2W_OT);
    p.ssb = s_stride_b; p.ssc = s_stride_c; p.ssp = s_stride_p;
    if (out_layout == HSB_W_PATCH_MAJOR) {
        HSB_REQUIRE(out_row_stride >= hp, HSB_ERR_INVALID_ARG, "signal2weights: out_row_stride < hp");
        p.osb = (int64_t)p.P * out_row_stride; p.osp = out_row_stride; p.osk = 1;
    } else {
        p.osb = (int64_t)hp * p.P; p.osp = 1; p.osk = p.P;
    }
    const int64_t ntiles = ((int64_t)B * p.P + S2W_NP - 1) / S2W_NP;
    HSB_REQUIRE(ntiles < (1ll << 31) && (int64_t)groups * p.otiles <= 65535, HSB_ERR_UNSUPPORTED,
                "signal2weights: grid too large");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)ntiles, groups * p.otiles);
    if (dtype == HSB_F32) signal2weights_kernel<float><<<grid, 256, 0, st>>>(p);
    else signal2weights_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(p);
    note_kernel("signal2weights_kernel");
    return check_launch("signal2weights launch");
}
