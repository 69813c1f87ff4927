// General patch-wise convolution (any k / groups / dilation), the standalone per-sample dynamic
// convolution, and the NCHW -> patch-major weight transposer.
//
// These serve the call sites that are not bandwidth-critical in the shipped models (HyperSeg-L's
// MetaPatchConv2d stages, `with_out_fc`) -- the 1x1 and inverted-residual blocks have their own
// kernels in patch_conv1x1.cu / patch_ir.cu.
//
// Semantics follow hyperseg/models/layers/meta_patch.py:35-57 + meta_conv.py:163-186 of the reference:
// pad the whole map, cut (ph+2*pad) x (pw+2*pad) tiles at stride (ph,pw), valid conv per tile.
#include "common.cuh"

namespace hsb {

struct PatchConvParams {
    const void* x; const void* w; void* y;
    const float* post_scale; const float* post_shift; int act;
    int B, Cin, Cout, H, W, fh, fw, ph, pw;
    int kh, kw, pad_h, pad_w, dil_h, dil_w, groups, pad_mode;
    WStrides ws;
    int cig, cog;     // channels per group (in / out)
    int th, tw;       // tile extent incl. halo
};

// One CTA per (patch, group). Shared memory: input tile of the group's Cin/G channels as fp32,
// followed by the group's weights as fp32.
template <typename T>
__global__ void __launch_bounds__(256) patch_conv_kernel(const PatchConvParams p) {
    extern __shared__ float smem[];
    const int P = p.fh * p.fw;
    const int patch = blockIdx.x;           // b*P + i*fw + j
    const int g = blockIdx.y;
    const int b = patch / P, pp = patch % P;
    const int pi = pp / p.fw, pj = pp % p.fw;
    const int tile_px = p.th * p.tw;
    float* tile = smem;                                   // [cig][th][tw]
    float* wsm = smem + (size_t)p.cig * tile_px;          // [cog][cig*kh*kw]
    const int kk = p.kh * p.kw;
    const int wpo = p.cig * kk;                           // weights per output channel

    const T* x = reinterpret_cast<const T*>(p.x);
    const T* w = reinterpret_cast<const T*>(p.w);
    T* y = reinterpret_cast<T*>(p.y);

    // stage the padded input tile
    const int y0 = pi * p.ph - p.pad_h, x0 = pj * p.pw - p.pad_w;
    for (int idx = threadIdx.x; idx < p.cig * tile_px; idx += blockDim.x) {
        int c = idx / tile_px, r = (idx % tile_px) / p.tw, q = idx % p.tw;
        bool vy, vx;
        int sy = pad_index(y0 + r, p.H, p.pad_mode, vy);
        int sx = pad_index(x0 + q, p.W, p.pad_mode, vx);
        float v = 0.f;
        if (vy && vx) v = ld_f(x + (((size_t)b * p.Cin + g * p.cig + c) * p.H + sy) * p.W + sx);
        tile[idx] = v;
    }
    // stage this group's weights
    const T* wp = w + (size_t)b * p.ws.b + (size_t)pp * p.ws.p;
    for (int idx = threadIdx.x; idx < p.cog * wpo; idx += blockDim.x) {
        size_t k = (size_t)(g * p.cog) * wpo + idx;
        wsm[idx] = ld_f(wp + k * p.ws.k);
    }
    __syncthreads();

    const int out_px = p.ph * p.pw;
    for (int idx = threadIdx.x; idx < p.cog * out_px; idx += blockDim.x) {
        int ol = idx / out_px, u = (idx % out_px) / p.pw, v = idx % p.pw;
        const float* wrow = wsm + (size_t)ol * wpo;
        float acc = 0.f;
        for (int c = 0; c < p.cig; ++c) {
            const float* tc = tile + (size_t)c * tile_px;
            for (int ky = 0; ky < p.kh; ++ky) {
                const float* tr = tc + (u + ky * p.dil_h) * p.tw + v;
                for (int kx = 0; kx < p.kw; ++kx)
                    acc = fmaf(wrow[(c * p.kh + ky) * p.kw + kx], tr[kx * p.dil_w], acc);
            }
        }
        int o = g * p.cog + ol;
        if (p.post_scale) acc = fmaf(acc, p.post_scale[o], p.post_shift[o]);
        acc = act_apply(acc, p.act);
        st_f(y + (((size_t)b * p.Cout + o) * p.H + pi * p.ph + u) * p.W + pj * p.pw + v, acc);
    }
}

// ---- standalone per-sample dynamic convolution (MetaConv2d) ---------------------------------
struct MetaConvParams {
    const void* x; const void* w; void* y;
    int N, Cin, Cout, H, W, Ho, Wo, kh, kw, pad_h, pad_w, dil_h, dil_w, groups, pad_mode;
};

template <typename T>
__global__ void __launch_bounds__(256) meta_conv2d_kernel(const MetaConvParams p) {
    const int cig = p.Cin / p.groups, cog = p.Cout / p.groups;
    const size_t total = (size_t)p.N * p.Cout * p.Ho * p.Wo;
    const T* x = reinterpret_cast<const T*>(p.x);
    const T* w = reinterpret_cast<const T*>(p.w);
    T* y = reinterpret_cast<T*>(p.y);
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        int v = idx % p.Wo;
        int u = (idx / p.Wo) % p.Ho;
        int o = (idx / ((size_t)p.Wo * p.Ho)) % p.Cout;
        int n = idx / ((size_t)p.Wo * p.Ho * p.Cout);
        int g = o / cog;
        const T* wn = w + ((size_t)n * p.Cout + o) * cig * p.kh * p.kw;
        float acc = 0.f;
        for (int c = 0; c < cig; ++c)
            for (int ky = 0; ky < p.kh; ++ky)
                for (int kx = 0; kx < p.kw; ++kx) {
                    bool vy, vx;
                    int sy = pad_index(u - p.pad_h + ky * p.dil_h, p.H, p.pad_mode, vy);
                    int sx = pad_index(v - p.pad_w + kx * p.dil_w, p.W, p.pad_mode, vx);
                    if (vy && vx)
                        acc = fmaf(ld_f(wn + (c * p.kh + ky) * p.kw + kx),
                                   ld_f(x + (((size_t)n * p.Cin + g * cig + c) * p.H + sy) * p.W + sx), acc);
                }
        st_f(y + idx, acc);
    }
}

// ---- (B, hp, P) -> (B, P, row_stride) ---------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) transpose_w_kernel(const T* __restrict__ in, T* __restrict__ out,
                                                          int hp, int P, int64_t row_stride) {
    __shared__ T tile[32][33];
    const int b = blockIdx.z;
    const int k0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
    const T* inb = in + (size_t)b * hp * P;
    T* outb = out + (size_t)b * P * row_stride;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int k = k0 + r, pp = p0 + threadIdx.x;
        if (k < hp && pp < P) tile[r][threadIdx.x] = inb[(size_t)k * P + pp];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int pp = p0 + r, k = k0 + threadIdx.x;
        if (k < hp && pp < P) outb[(size_t)pp * row_stride + k] = tile[threadIdx.x][r];
    }
}

}  // namespace hsb

using namespace hsb;

extern "C" int hsb_patch_conv_fwd(const void* x, const void* w, void* y,
                                  const float* post_scale, const float* post_shift, int act,
                                  int B, int Cin, int Cout, int H, int W, int fh, int fw,
                                  int kh, int kw, int pad_h, int pad_w, int dil_h, int dil_w, int groups,
                                  int pad_mode, int dtype, int w_layout, int64_t w_row_stride, void* stream) {
    HSB_REQUIRE(x && w && y, HSB_ERR_INVALID_ARG, "patch_conv: null pointer");
    HSB_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && H > 0 && W > 0 && fh > 0 && fw > 0 && groups > 0,
                HSB_ERR_INVALID_ARG, "patch_conv: non-positive dimension");
    HSB_REQUIRE(H % fh == 0 && W % fw == 0, HSB_ERR_INVALID_ARG,
                "patch_conv: feature map is not divisible into fh x fw patches");
    HSB_REQUIRE(Cin % groups == 0 && Cout % groups == 0, HSB_ERR_INVALID_ARG,
                "patch_conv: channels not divisible by groups");
    HSB_REQUIRE((post_scale == nullptr) == (post_shift == nullptr), HSB_ERR_INVALID_ARG,
                "patch_conv: post_scale and post_shift must be given together");
    HSB_REQUIRE(kh > 0 && kw > 0 && dil_h > 0 && dil_w > 0 && pad_h >= 0 && pad_w >= 0,
                HSB_ERR_INVALID_ARG, "patch_conv: bad kernel geometry");
    HSB_REQUIRE(2 * pad_h == dil_h * (kh - 1) && 2 * pad_w == dil_w * (kw - 1), HSB_ERR_UNSUPPORTED,
                "patch_conv: only size-preserving geometry (2*pad == dilation*(k-1)) is supported");
    HSB_REQUIRE(dtype == HSB_F32 || dtype == HSB_BF16, HSB_ERR_INVALID_ARG, "patch_conv: bad dtype");
    if (pad_mode == HSB_PAD_REFLECT)
        HSB_REQUIRE(pad_h < H && pad_w < W, HSB_ERR_INVALID_ARG, "patch_conv: reflect pad >= size");
    PatchConvParams p;
    p.x = x; p.w = w; p.y = y; p.post_scale = post_scale; p.post_shift = post_shift; p.act = act;
    p.B = B; p.Cin = Cin; p.Cout = Cout; p.H = H; p.W = W; p.fh = fh; p.fw = fw;
    p.ph = H / fh; p.pw = W / fw;
    p.kh = kh; p.kw = kw; p.pad_h = pad_h; p.pad_w = pad_w; p.dil_h = dil_h; p.dil_w = dil_w;
    p.groups = groups; p.pad_mode = pad_mode;
    p.cig = Cin / groups; p.cog = Cout / groups;
    p.th = p.ph + 2 * pad_h; p.tw = p.pw + 2 * pad_w;
    const int64_t hp = (int64_t)Cout * p.cig * kh * kw;
    if (w_layout == HSB_W_PATCH_MAJOR)
        HSB_REQUIRE(w_row_stride >= hp, HSB_ERR_INVALID_ARG, "patch_conv: w_row_stride < hyper params");
    p.ws = make_wstrides(w_layout, hp, (int64_t)fh * fw, w_row_stride);
    size_t smem = ((size_t)p.cig * p.th * p.tw + (size_t)p.cog * p.cig * kh * kw) * sizeof(float);
    HSB_REQUIRE(smem <= 220 * 1024, HSB_ERR_UNSUPPORTED,
                "patch_conv: patch tile + weights exceed shared memory (" + std::to_string(smem) + " B)");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid(B * fh * fw, groups);
    HSB_REQUIRE(groups <= 65535, HSB_ERR_UNSUPPORTED, "patch_conv: groups > 65535");
    cudaError_t e;
    if (dtype == HSB_F32) {
        e = cudaFuncSetAttribute(patch_conv_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(HSB_ERR_CUDA, std::string("patch_conv attr: ") + cudaGetErrorString(e));
        patch_conv_kernel<float><<<grid, 256, smem, st>>>(p);
    } else {
        e = cudaFuncSetAttribute(patch_conv_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(HSB_ERR_CUDA, std::string("patch_conv attr: ") + cudaGetErrorString(e));
        patch_conv_kernel<__nv_bfloat16><<<grid, 256, smem, st>>>(p);
    }
    note_kernel("patch_conv_kernel");
    return check_launch("patch_conv launch");
}

extern "C" int hsb_meta_conv2d_fwd(const void* x, const void* w, void* y,
                                   int N, int Cin, int Cout, int H, int W,
                                   int kh, int kw, int pad_h, int pad_w, int dil_h, int dil_w, int groups,
                                   int pad_mode, int dtype, void* stream) {
    HSB_REQUIRE(x && w && y, HSB_ERR_INVALID_ARG, "meta_conv2d: null pointer");
    HSB_REQUIRE(N > 0 && Cin > 0 && Cout > 0 && H > 0 && W > 0 && groups > 0 && kh > 0 && kw > 0 &&
                dil_h > 0 && dil_w > 0 && pad_h >= 0 && pad_w >= 0,
                HSB_ERR_INVALID_ARG, "meta_conv2d: bad dimension");
    HSB_REQUIRE(Cin % groups == 0 && Cout % groups == 0, HSB_ERR_INVALID_ARG,
                "meta_conv2d: channels not divisible by groups");
    HSB_REQUIRE(dtype == HSB_F32 || dtype == HSB_BF16, HSB_ERR_INVALID_ARG, "meta_conv2d: bad dtype");
    MetaConvParams p;
    p.x = x; p.w = w; p.y = y; p.N = N; p.Cin = Cin; p.Cout = Cout; p.H = H; p.W = W;
    p.kh = kh; p.kw = kw; p.pad_h = pad_h; p.pad_w = pad_w; p.dil_h = dil_h; p.dil_w = dil_w;
    p.groups = groups; p.pad_mode = pad_mode;
    p.Ho = H + 2 * pad_h - dil_h * (kh - 1);
    p.Wo = W + 2 * pad_w - dil_w * (kw - 1);
    HSB_REQUIRE(p.Ho > 0 && p.Wo > 0, HSB_ERR_INVALID_ARG, "meta_conv2d: empty output");
    if (pad_mode == HSB_PAD_REFLECT)
        HSB_REQUIRE(pad_h < H && pad_w < W, HSB_ERR_INVALID_ARG, "meta_conv2d: reflect pad >= size");
    size_t total = (size_t)N * Cout * p.Ho * p.Wo;
    int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)device_sm_count() * 16);
    if (blocks < 1) blocks = 1;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == HSB_F32) meta_conv2d_kernel<float><<<blocks, 256, 0, st>>>(p);
    else meta_conv2d_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(p);
    return check_launch("meta_conv2d launch");
}

extern "C" int hsb_weights_to_patch_major(const void* w_nchw, void* w_pm, int B, int hp, int fh, int fw,
                                          int64_t row_stride, int dtype, void* stream) {
    HSB_REQUIRE(w_nchw && w_pm, HSB_ERR_INVALID_ARG, "weights_to_patch_major: null pointer");
    HSB_REQUIRE(B > 0 && hp > 0 && fh > 0 && fw > 0 && row_stride >= hp, HSB_ERR_INVALID_ARG,
                "weights_to_patch_major: bad dimension");
    HSB_REQUIRE(dtype == HSB_F32 || dtype == HSB_BF16, HSB_ERR_INVALID_ARG, "weights_to_patch_major: bad dtype");
    HSB_REQUIRE(B <= 65535, HSB_ERR_UNSUPPORTED, "weights_to_patch_major: B > 65535");
    const int P = fh * fw;
    dim3 grid(ceil_div(P, 32), ceil_div(hp, 32), B), block(32, 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == HSB_F32)
        transpose_w_kernel<float><<<grid, block, 0, st>>>((const float*)w_nchw, (float*)w_pm, hp, P, row_stride);
    else
        transpose_w_kernel<__nv_bfloat16><<<grid, block, 0, st>>>((const __nv_bfloat16*)w_nchw,
                                                                  (__nv_bfloat16*)w_pm, hp, P, row_stride);
    return check_launch("weights_to_patch_major launch");
}
