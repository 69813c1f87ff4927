// Fused patch-wise inverted-residual MetaBlock -- CUDA-core fp32 path (any shape, fp32 or bf16 I/O).
//
// Replaces HyperPatchInvertedResidual.conv/forward (reference hyperseg/models/hyperseg_v1_0.py:328-376):
// reflect pad + unfold + 3 grouped convs + 3 BatchNorms + 2 ReLU6 + re-tile become one kernel; the
// (ph+2)x(pw+2) halo tile, the expanded hidden tensor and the depthwise output never leave the SM.
//
// One CTA per patch.  The hidden dimension is processed in chunks of HC channels so that the fp32
// staging buffers of the largest shipped shape (Cin 34, hid 68, 18x18 tile) fit twice per SM:
//   A  h[hc][t]   = relu6(s1 * sum_c W1[hc][c] * tile[c][t] + b1)      (all (ph+2)(pw+2) positions)
//   B  d[hc][u,v] = relu6(s2 * sum_{3x3} W2[hc][ky,kx] * h[hc][u+ky,v+kx] + b2)
//   C  acc[o]    += sum_{hc in chunk} W3[o][hc] * d[hc][u,v]               (registers, thread = pixel)
// and after the last chunk  y = s3*acc + b3 (+x).
// This is the exact-parity path (fp32 accumulate everywhere); the bf16 tensor-core path lives in
// patch_ir_tc.cu and is preferred by the dispatcher whenever its shape constraints hold.
#include "common.cuh"

namespace hsb {

int launch_patch_ir_tc(const void* x, const void* w, void* y, const float* const* bn, int B, int Cin, int hid,
                       int Cout, int H, int W, int fh, int fw, int residual, int64_t w_row_stride,
                       cudaStream_t st, bool* handled);

struct IRParams {
    const void* x; const void* w; void* y;
    const float* s1; const float* b1; const float* s2; const float* b2; const float* s3; const float* b3;
    int B, Cin, hid, Cout, H, W, fh, fw, ph, pw, th, tw, residual;
    WStrides ws;
    int hp, HC;
};

template <typename T, int CO_MAX, int NPX>
__global__ void __launch_bounds__(256) patch_ir_kernel(const IRParams p) {
    extern __shared__ float smem[];
    const int T_ = p.th * p.tw, O_ = p.ph * p.pw;
    float* xs = smem;                              // [Cin][T_]
    float* wsm = xs + (size_t)p.Cin * T_;          // [hp]
    float* hs = wsm + p.hp;                        // [HC][T_]
    float* ds = hs + (size_t)p.HC * T_;            // [HC][O_]
    const int tid = threadIdx.x;
    const int P = p.fh * p.fw;
    const int b = blockIdx.x / P, pp = blockIdx.x % P;
    const int pi = pp / p.fw, pj = pp % p.fw;
    const T* x = reinterpret_cast<const T*>(p.x);
    const T* w = reinterpret_cast<const T*>(p.w);
    T* y = reinterpret_cast<T*>(p.y);

    // halo tile with reflect padding at the image border (interior halos are real neighbours)
    const int y0 = pi * p.ph - 1, x0 = pj * p.pw - 1;
    for (int idx = tid; idx < p.Cin * T_; idx += blockDim.x) {
        int c = idx / T_, r = (idx % T_) / p.tw, q = idx % p.tw;
        bool v1, v2;
        int sy = pad_index(y0 + r, p.H, HSB_PAD_REFLECT, v1);
        int sx = pad_index(x0 + q, p.W, HSB_PAD_REFLECT, v2);
        xs[idx] = ld_f(x + (((size_t)b * p.Cin + c) * p.H + sy) * p.W + sx);
    }
    const T* wp = w + (size_t)b * p.ws.b + (size_t)pp * p.ws.p;
    for (int k = tid; k < p.hp; k += blockDim.x) wsm[k] = ld_f(wp + (size_t)k * p.ws.k);
    __syncthreads();

    const float* W1 = wsm;                               // [hid][Cin]
    const float* W2 = wsm + (size_t)p.hid * p.Cin;       // [hid][9]
    const float* W3 = W2 + (size_t)p.hid * 9;            // [Cout][hid]

    float acc[NPX][CO_MAX];
#pragma unroll
    for (int n = 0; n < NPX; ++n)
#pragma unroll
        for (int o = 0; o < CO_MAX; ++o) acc[n][o] = 0.f;

    for (int h0 = 0; h0 < p.hid; h0 += p.HC) {
        const int hc_n = min(p.HC, p.hid - h0);
        // ---- A: pointwise expansion on the halo tile, 4 hidden channels per work item ----
        const int hgroups = (hc_n + 3) / 4;
        for (int it = tid; it < hgroups * T_; it += blockDim.x) {
            const int t = it % T_, hg = it / T_;
            const int hl = hg * 4;                       // local hidden index
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            const float* w0 = W1 + (size_t)(h0 + hl) * p.Cin;
            const bool k1 = hl + 1 < hc_n, k2 = hl + 2 < hc_n, k3 = hl + 3 < hc_n;
            const float* w1 = k1 ? w0 + p.Cin : w0;
            const float* w2 = k2 ? w0 + 2 * p.Cin : w0;
            const float* w3 = k3 ? w0 + 3 * p.Cin : w0;
            for (int c = 0; c < p.Cin; ++c) {
                float xv = xs[(size_t)c * T_ + t];
                a0 = fmaf(w0[c], xv, a0);
                a1 = fmaf(w1[c], xv, a1);
                a2 = fmaf(w2[c], xv, a2);
                a3 = fmaf(w3[c], xv, a3);
            }
            const int hgl = h0 + hl;
            hs[(size_t)hl * T_ + t] = fminf(fmaxf(fmaf(a0, p.s1[hgl], p.b1[hgl]), 0.f), 6.f);
            if (k1) hs[(size_t)(hl + 1) * T_ + t] = fminf(fmaxf(fmaf(a1, p.s1[hgl + 1], p.b1[hgl + 1]), 0.f), 6.f);
            if (k2) hs[(size_t)(hl + 2) * T_ + t] = fminf(fmaxf(fmaf(a2, p.s1[hgl + 2], p.b1[hgl + 2]), 0.f), 6.f);
            if (k3) hs[(size_t)(hl + 3) * T_ + t] = fminf(fmaxf(fmaf(a3, p.s1[hgl + 3], p.b1[hgl + 3]), 0.f), 6.f);
        }
        __syncthreads();
        // ---- B: depthwise 3x3 (valid) ----
        for (int it = tid; it < hc_n * O_; it += blockDim.x) {
            const int px = it % O_, hl = it / O_;
            const int u = px / p.pw, v = px % p.pw;
            const float* k9 = W2 + (size_t)(h0 + hl) * 9;
            const float* hrow = hs + (size_t)hl * T_ + u * p.tw + v;
            float a = 0.f;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) a = fmaf(k9[ky * 3 + kx], hrow[ky * p.tw + kx], a);
            const int hgl = h0 + hl;
            ds[(size_t)hl * O_ + px] = fminf(fmaxf(fmaf(a, p.s2[hgl], p.b2[hgl]), 0.f), 6.f);
        }
        __syncthreads();
        // ---- C: pointwise projection, accumulated across chunks ----
#pragma unroll
        for (int n = 0; n < NPX; ++n) {
            const int px = tid + n * 256;
            if (px < O_) {
                for (int hl = 0; hl < hc_n; ++hl) {
                    const float dv = ds[(size_t)hl * O_ + px];
                    const float* w3c = W3 + (h0 + hl);
#pragma unroll
                    for (int o = 0; o < CO_MAX; ++o)
                        if (o < p.Cout) acc[n][o] = fmaf(w3c[(size_t)o * p.hid], dv, acc[n][o]);
                }
            }
        }
        __syncthreads();   // hs / ds are rewritten by the next chunk
    }

#pragma unroll
    for (int n = 0; n < NPX; ++n) {
        const int px = tid + n * 256;
        if (px < O_) {
            const int u = px / p.pw, v = px % p.pw;
            const size_t pix = ((size_t)pi * p.ph + u) * p.W + (size_t)pj * p.pw + v;
#pragma unroll
            for (int o = 0; o < CO_MAX; ++o) {
                if (o < p.Cout) {
                    float r = fmaf(acc[n][o], p.s3[o], p.b3[o]);
                    if (p.residual) r += xs[(size_t)o * T_ + (u + 1) * p.tw + v + 1];
                    st_f(y + ((size_t)b * p.Cout + o) * p.H * p.W + pix, r);
                }
            }
        }
    }
}

template <typename T, int CO_MAX, int NPX>
static int launch_ir(const IRParams& p, size_t smem, cudaStream_t st) {
    auto k = patch_ir_kernel<T, CO_MAX, NPX>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(HSB_ERR_CUDA, std::string("patch_ir attr: ") + cudaGetErrorString(e));
    k<<<p.B * p.fh * p.fw, 256, smem, st>>>(p);
    note_kernel("patch_ir_kernel");
    return check_launch("patch_ir launch");
}

template <typename T>
static int dispatch_ir(IRParams& p, cudaStream_t st) {
    const int T_ = p.th * p.tw, O_ = p.ph * p.pw;
    const size_t fixed = (size_t)p.Cin * T_ + p.hp;
    const size_t per_h = (size_t)T_ + O_;
    const size_t budget_2cta = 110 * 1024 / sizeof(float), budget_max = 220 * 1024 / sizeof(float);
    HSB_REQUIRE(fixed + 4 * per_h <= budget_max, HSB_ERR_UNSUPPORTED,
                "patch_ir: patch tile too large for shared memory");
    int HC;
    if (fixed + (size_t)p.hid * per_h <= budget_2cta) HC = p.hid;
    else {
        size_t budget = fixed + 8 * per_h <= budget_2cta ? budget_2cta : budget_max;
        HC = (int)((budget - fixed) / per_h);
        HC = std::max(4, HC / 4 * 4);
        HC = std::min(HC, (p.hid + 3) / 4 * 4);
        // balance the chunks
        int nch = ceil_div(p.hid, HC);
        HC = std::min(HC, (ceil_div(p.hid, nch) + 3) / 4 * 4);
    }
    p.HC = HC;
    const size_t smem = (fixed + (size_t)HC * per_h) * sizeof(float);
    const int npx = ceil_div(O_, 256);
    HSB_REQUIRE(npx <= 4, HSB_ERR_UNSUPPORTED, "patch_ir: patches larger than 1024 pixels are not supported");
    HSB_REQUIRE(p.Cout <= 64, HSB_ERR_UNSUPPORTED, "patch_ir: more than 64 output channels is not supported");
    HSB_REQUIRE(p.Cout <= 32 || npx == 1, HSB_ERR_UNSUPPORTED,
                "patch_ir: >32 output channels needs patches of at most 256 pixels");
#define HSB_IR_CASE(CO, NP) return launch_ir<T, CO, NP>(p, smem, st)
    if (p.Cout <= 8) { if (npx == 1) HSB_IR_CASE(8, 1); if (npx == 2) HSB_IR_CASE(8, 2); HSB_IR_CASE(8, 4); }
    if (p.Cout <= 16) { if (npx == 1) HSB_IR_CASE(16, 1); if (npx == 2) HSB_IR_CASE(16, 2); HSB_IR_CASE(16, 4); }
    if (p.Cout <= 24) { if (npx == 1) HSB_IR_CASE(24, 1); if (npx == 2) HSB_IR_CASE(24, 2); HSB_IR_CASE(24, 4); }
    if (p.Cout <= 32) { if (npx == 1) HSB_IR_CASE(32, 1); if (npx == 2) HSB_IR_CASE(32, 2); HSB_IR_CASE(32, 4); }
    HSB_IR_CASE(64, 1);
#undef HSB_IR_CASE
}

}  // namespace hsb

using namespace hsb;

extern "C" int hsb_patch_ir_fwd(const void* x, const void* w, void* y,
                                const float* bn1_scale, const float* bn1_shift,
                                const float* bn2_scale, const float* bn2_shift,
                                const float* bn3_scale, const float* bn3_shift,
                                int B, int Cin, int hid, int Cout, int H, int W, int fh, int fw,
                                int residual, int dtype, int w_layout, int64_t w_row_stride, void* stream) {
    HSB_REQUIRE(x && w && y, HSB_ERR_INVALID_ARG, "patch_ir: null pointer");
    HSB_REQUIRE(bn1_scale && bn1_shift && bn2_scale && bn2_shift && bn3_scale && bn3_shift,
                HSB_ERR_INVALID_ARG, "patch_ir: null BatchNorm scale/shift");
    HSB_REQUIRE(B > 0 && Cin > 0 && hid > 0 && Cout > 0 && H > 0 && W > 0 && fh > 0 && fw > 0,
                HSB_ERR_INVALID_ARG, "patch_ir: non-positive dimension");
    HSB_REQUIRE(H % fh == 0 && W % fw == 0, HSB_ERR_INVALID_ARG,
                "patch_ir: feature map is not divisible into fh x fw patches");
    HSB_REQUIRE(H >= 2 && W >= 2, HSB_ERR_INVALID_ARG, "patch_ir: reflect padding needs H,W >= 2");
    HSB_REQUIRE(!residual || Cin == Cout, HSB_ERR_INVALID_ARG, "patch_ir: residual needs Cin == Cout");
    HSB_REQUIRE(dtype == HSB_F32 || dtype == HSB_BF16, HSB_ERR_INVALID_ARG, "patch_ir: bad dtype");
    HSB_REQUIRE((int64_t)B * fh * fw < (1ll << 31), HSB_ERR_UNSUPPORTED, "patch_ir: too many patches");
    IRParams p;
    p.x = x; p.w = w; p.y = y;
    p.s1 = bn1_scale; p.b1 = bn1_shift; p.s2 = bn2_scale; p.b2 = bn2_shift; p.s3 = bn3_scale; p.b3 = bn3_shift;
    p.B = B; p.Cin = Cin; p.hid = hid; p.Cout = Cout; p.H = H; p.W = W; p.fh = fh; p.fw = fw;
    p.ph = H / fh; p.pw = W / fw; p.th = p.ph + 2; p.tw = p.pw + 2; p.residual = residual ? 1 : 0;
    p.hp = Cin * hid + 9 * hid + hid * Cout;
    if (w_layout == HSB_W_PATCH_MAJOR)
        HSB_REQUIRE(w_row_stride >= p.hp, HSB_ERR_INVALID_ARG, "patch_ir: w_row_stride < hyper params");
    p.ws = make_wstrides(w_layout, p.hp, (int64_t)fh * fw, w_row_stride);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == HSB_BF16 && w_layout == HSB_W_PATCH_MAJOR) {
        const float* bn[6] = {bn1_scale, bn1_shift, bn2_scale, bn2_shift, bn3_scale, bn3_shift};
        bool handled = false;
        int rc = launch_patch_ir_tc(x, w, y, bn, B, Cin, hid, Cout, H, W, fh, fw, p.residual, w_row_stride, st,
                                    &handled);
        if (handled) return rc;
    }
    if (dtype == HSB_F32) return dispatch_ir<float>(p, st);
    return dispatch_ir<__nv_bfloat16>(p, st);
}
