// Backward of the general patch-wise convolution and of the weight head (training path).
//
// Forward (patch_generic.cu; reference hyperseg/models/layers/meta_patch.py:35-57 + meta_conv.py:163-186):
//     xp = pad(x);  y[b,o,oy,ox] = sum_{c,ky,kx} Wm[b,patch(oy,ox)][o,c,ky,kx] * xp[b, g*cig+c, oy + ky*d, ox + kx*d]
// In the reference the gradients come from autograd through F.pad / F.unfold / F.conv2d(groups) / F.fold; here
//     hsb_patch_conv_bwd_weight   dW[b,patch][o,c,ky,kx] = sum_{pixels of the patch} dy * xp            (one CTA per patch, group)
//     hsb_patch_conv_bwd_input    dx[b,ci,y,x]  = sum over the padded coordinates that alias (y,x), the taps and the output
//                                 channels of  dy[b,o,oy,ox] * Wm[b,patch(oy,ox)][o,c,ky,kx]              (gather: no atomics)
//     hsb_signal2weights_bwd_signal / _bwd_weight   the two gradients of the grouped 1x1 head (hyperseg_v1_0.py:315-326).
// These are correctness-first CUDA-core kernels (fp32 accumulation, fp32 or bf16 storage); the forward kernels are the
// optimised path.  Used for BASELINE config 4 (HyperSeg-L / hyperseg_v0_1 training step).
#include "common.cuh"

namespace hsb {

struct PatchBwdParams {
    const void* x; const void* w; const void* dy; void* dx; void* dw;
    int B, Cin, Cout, H, W, fh, fw, ph, pw;
    int kh, kw, pad_h, pad_w, dil_h, dil_w, groups, pad_mode;
    WStrides ws;      // strides of w (read) and of dw (written): same layout
    int cig, cog, th, tw;
};

// ---- dW: one CTA per (patch, group); padded input tile + dy tile in shared memory ----------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) patch_conv_bwd_weight_kernel(const PatchBwdParams p) {
    extern __shared__ float smem[];
    const int P = p.fh * p.fw;
    const int patch = blockIdx.x, g = blockIdx.y;
    const int b = patch / P, pp = patch % P, pi = pp / p.fw, pj = pp % p.fw;
    const int tile_px = p.th * p.tw, out_px = p.ph * p.pw;
    float* tile = smem;                                  // [cig][th][tw]
    float* dys = smem + (size_t)p.cig * tile_px;         // [cog][ph][pw]
    const T* x = reinterpret_cast<const T*>(p.x);
    const T* dy = reinterpret_cast<const T*>(p.dy);
    T* dw = reinterpret_cast<T*>(p.dw);
    const int y0 = pi * p.ph - p.pad_h, x0 = pj * p.pw - p.pad_w;
    for (int idx = threadIdx.x; idx < p.cig * tile_px; idx += blockDim.x) {
        int c = idx / tile_px, r = (idx % tile_px) / p.tw, q = idx % p.tw;
        bool vy, vx;
        int sy = pad_index(y0 + r, p.H, p.pad_mode, vy), sx = pad_index(x0 + q, p.W, p.pad_mode, vx);
        tile[idx] = (vy && vx) ? ld_f(x + (((size_t)b * p.Cin + g * p.cig + c) * p.H + sy) * p.W + sx) : 0.f;
    }
    for (int idx = threadIdx.x; idx < p.cog * out_px; idx += blockDim.x) {
        int ol = idx / out_px, u = (idx % out_px) / p.pw, v = idx % p.pw;
        dys[idx] = ld_f(dy + (((size_t)b * p.Cout + g * p.cog + ol) * p.H + pi * p.ph + u) * p.W + pj * p.pw + v);
    }
    __syncthreads();
    const int kk = p.kh * p.kw, wpo = p.cig * kk;
    T* dwp = dw + (size_t)b * p.ws.b + (size_t)pp * p.ws.p;
    for (int idx = threadIdx.x; idx < p.cog * wpo; idx += blockDim.x) {
        const int ol = idx / wpo, c = (idx % wpo) / kk, ky = (idx % kk) / p.kw, kx = idx % p.kw;
        const float* tc = tile + (size_t)c * tile_px + (ky * p.dil_h) * p.tw + kx * p.dil_w;
        const float* dr = dys + (size_t)ol * out_px;
        float acc = 0.f;
        for (int u = 0; u < p.ph; ++u)
            for (int v = 0; v < p.pw; ++v) acc = fmaf(dr[u * p.pw + v], tc[u * p.tw + v], acc);
        st_f(dwp + ((size_t)(g * p.cog) * wpo + idx) * p.ws.k, acc);
    }
}

// padded coordinates t (relative to the unpadded map, t in [-pad, n+pad)) whose source sample is `i`
__device__ __forceinline__ int pad_aliases(int i, int n, int pad, int mode, int (&t)[4]) {
    int cnt = 0;
    t[cnt++] = i;
    if (pad == 0) return cnt;
    if (mode == HSB_PAD_REFLECT) {
        if (i >= 1 && i <= pad) t[cnt++] = -i;
        if (i <= n - 2 && i >= n - 1 - pad) t[cnt++] = 2 * (n - 1) - i;
    } else if (mode == HSB_PAD_REPLICATE) {
        if (i == 0) for (int k = 1; k <= pad && cnt < 4; ++k) t[cnt++] = -k;
        if (i == n - 1) for (int k = 0; k < pad && cnt < 4; ++k) t[cnt++] = n + k;
    } else if (mode == HSB_PAD_CIRCULAR) {
        if (i >= n - pad) t[cnt++] = i - n;
        if (i < pad) t[cnt++] = i + n;
    }
    return cnt;
}

// ---- dx: one thread per input element, gathering from every output pixel that read it ------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) patch_conv_bwd_input_kernel(const PatchBwdParams p) {
    const size_t total = (size_t)p.B * p.Cin * p.H * p.W;
    const T* w = reinterpret_cast<const T*>(p.w);
    const T* dy = reinterpret_cast<const T*>(p.dy);
    T* dx = reinterpret_cast<T*>(p.dx);
    const int kk = p.kh * p.kw, wpo = p.cig * kk;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int xq = idx % p.W, yq = (idx / p.W) % p.H;
        const int ci = (idx / ((size_t)p.W * p.H)) % p.Cin, b = idx / ((size_t)p.W * p.H * p.Cin);
        const int g = ci / p.cig, c = ci % p.cig;
        int ty[4], tx[4];
        const int ny = pad_aliases(yq, p.H, p.pad_h, p.pad_mode, ty), nx = pad_aliases(xq, p.W, p.pad_w, p.pad_mode, tx);
        float acc = 0.f;
        for (int a = 0; a < ny; ++a)
            for (int ky = 0; ky < p.kh; ++ky) {
                const int oy = ty[a] + p.pad_h - ky * p.dil_h;          // output row that read padded row ty[a] with tap ky
                if (oy < 0 || oy >= p.H) continue;
                for (int e = 0; e < nx; ++e)
                    for (int kx = 0; kx < p.kw; ++kx) {
                        const int ox = tx[e] + p.pad_w - kx * p.dil_w;
                        if (ox < 0 || ox >= p.W) continue;
                        const int pp = (oy / p.ph) * p.fw + ox / p.pw;
                        const T* wp = w + (size_t)b * p.ws.b + (size_t)pp * p.ws.p;
                        const T* dyp = dy + (((size_t)b * p.Cout + g * p.cog) * p.H + oy) * p.W + ox;
                        for (int ol = 0; ol < p.cog; ++ol) {
                            const size_t k = (size_t)(g * p.cog + ol) * wpo + (c * p.kh + ky) * p.kw + kx;
                            acc = fmaf(ld_f(dyp + (size_t)ol * p.H * p.W), ld_f(wp + k * p.ws.k), acc);
                        }
                    }
            }
        st_f(dx + idx, acc);
    }
}

// ---- head backward --------------------------------------------------------------------------------------------------------
struct HeadBwdParams {
    const void* s; const void* ws; const void* dwout; void* ds; void* dws;
    int B, P, sig_total, sig_index, sig_ch, out_ch, hp, groups, spg, opg;
    int64_t ssb, ssc, ssp;       // strides of s (read) / ds (written, same layout, all sig_total channels)
    int64_t gsb, gsp, gsk;       // strides of dwout: image, patch, weight index
};

// ds[b, ch, n]: channels outside [sig_index, sig_index+sig_ch) get zero
template <typename T>
__global__ void __launch_bounds__(256) head_bwd_signal_kernel(const HeadBwdParams p) {
    const size_t total = (size_t)p.B * p.sig_total * p.P;
    const T* ws = reinterpret_cast<const T*>(p.ws);
    const T* dwo = reinterpret_cast<const T*>(p.dwout);
    T* ds = reinterpret_cast<T*>(p.ds);
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int pp = idx % p.P, ch = (idx / p.P) % p.sig_total, b = idx / ((size_t)p.P * p.sig_total);
        float acc = 0.f;
        const int rel = ch - p.sig_index;
        if (rel >= 0 && rel < p.sig_ch) {
            const int g = rel / p.spg, k = rel % p.spg;
            const int o0 = g * p.opg, o1 = min((g + 1) * p.opg, p.hp);
            const T* grow = dwo + (size_t)b * p.gsb + (size_t)pp * p.gsp;
            for (int o = o0; o < o1; ++o) acc = fmaf(ld_f(ws + (size_t)o * p.spg + k), ld_f(grow + (size_t)o * p.gsk), acc);
        }
        st_f(ds + (size_t)b * p.ssb + (size_t)ch * p.ssc + (size_t)pp * p.ssp, acc);
    }
}

// dWs[o, k] = sum_n dwout[n, o] * s[n, idx + g*K + k]; one warp per (o, k), lanes over positions
template <typename T>
__global__ void __launch_bounds__(256) head_bwd_weight_kernel(const HeadBwdParams p) {
    const int lane = threadIdx.x & 31;
    const size_t warp_id = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    const T* s = reinterpret_cast<const T*>(p.s);
    const T* dwo = reinterpret_cast<const T*>(p.dwout);
    T* dws = reinterpret_cast<T*>(p.dws);
    const size_t total = (size_t)p.out_ch * p.spg;
    const int N = p.B * p.P;
    for (size_t item = warp_id; item < total; item += nwarps) {
        const int o = item / p.spg, k = item % p.spg;
        float acc = 0.f;
        if (o < p.hp) {
            const int g = o / p.opg;
            const int ch = p.sig_index + g * p.spg + k;
            for (int n = lane; n < N; n += 32) {
                const int b = n / p.P, pp = n % p.P;
                acc = fmaf(ld_f(dwo + (size_t)b * p.gsb + (size_t)pp * p.gsp + (size_t)o * p.gsk),
                           ld_f(s + (size_t)b * p.ssb + (size_t)ch * p.ssc + (size_t)pp * p.ssp), acc);
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (lane == 0) st_f(dws + item, acc);          // rows o >= hp (group padding) get a zero gradient
    }
}

static int fill_patch_bwd(PatchBwdParams& p, int B, int Cin, int Cout, int H, int W, int fh, int fw, int kh, int kw,
                          int pad_h, int pad_w, int dil_h, int dil_w, int groups, int pad_mode, int dtype, int w_layout,
                          int64_t w_row_stride, const char* who) {
    HSB_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && H > 0 && W > 0 && fh > 0 && fw > 0 && groups > 0 && kh > 0 && kw > 0,
                HSB_ERR_INVALID_ARG, std::string(who) + ": non-positive dimension");
    HSB_REQUIRE(H % fh == 0 && W % fw == 0 && Cin % groups == 0 && Cout % groups == 0, HSB_ERR_INVALID_ARG,
                std::string(who) + ": bad patch grid / groups");
    HSB_REQUIRE(2 * pad_h == dil_h * (kh - 1) && 2 * pad_w == dil_w * (kw - 1), HSB_ERR_UNSUPPORTED,
                std::string(who) + ": only size-preserving geometry is supported");
    HSB_REQUIRE(pad_h <= 3 && pad_w <= 3, HSB_ERR_UNSUPPORTED, std::string(who) + ": padding > 3");
    HSB_REQUIRE(dtype == HSB_F32 || dtype == HSB_BF16, HSB_ERR_INVALID_ARG, std::string(who) + ": bad dtype");
    if (pad_mode == HSB_PAD_REFLECT)
        HSB_REQUIRE(pad_h < H && pad_w < W, HSB_ERR_INVALID_ARG, std::string(who) + ": reflect pad >= size");
    p.B = B; p.Cin = Cin; p.Cout = Cout; p.H = H; p.W = W; p.fh = fh; p.fw = fw; p.ph = H / fh; p.pw = W / fw;
    p.kh = kh; p.kw = kw; p.pad_h = pad_h; p.pad_w = pad_w; p.dil_h = dil_h; p.dil_w = dil_w; p.groups = groups;
    p.pad_mode = pad_mode; p.cig = Cin / groups; p.cog = Cout / groups; p.th = p.ph + 2 * pad_h; p.tw = p.pw + 2 * pad_w;
    const int64_t hp = (int64_t)Cout * p.cig * kh * kw;
    if (w_layout == HSB_W_PATCH_MAJOR)
        HSB_REQUIRE(w_row_stride >= hp, HSB_ERR_INVALID_ARG, std::string(who) + ": w_row_stride < hyper params");
    p.ws = make_wstrides(w_layout, hp, (int64_t)fh * fw, w_row_stride);
    return HSB_OK;
}

}  // namespace hsb

using namespace hsb;

extern "C" int hsb_patch_conv_bwd_weight(const void* x, const void* dy, void* dw,
                                         int B, int Cin, int Cout, int H, int W, int fh, int fw,
                                         int kh, int kw, int pad_h, int pad_w, int dil_h, int dil_w, int groups,
                                         int pad_mode, int dtype, int w_layout, int64_t w_row_stride, void* stream) {
    HSB_REQUIRE(x && dy && dw, HSB_ERR_INVALID_ARG, "patch_conv_bwd_weight: null pointer");
    PatchBwdParams p{};
    int rc = fill_patch_bwd(p, B, Cin, Cout, H, W, fh, fw, kh, kw, pad_h, pad_w, dil_h, dil_w, groups, pad_mode, dtype,
                            w_layout, w_row_stride, "patch_conv_bwd_weight");
    if (rc != HSB_OK) return rc;
    p.x = x; p.dy = dy; p.dw = dw;
    const size_t smem = ((size_t)p.cig * p.th * p.tw + (size_t)p.cog * p.ph * p.pw) * sizeof(float);
    HSB_REQUIRE(smem <= 220 * 1024, HSB_ERR_UNSUPPORTED, "patch_conv_bwd_weight: patch tiles exceed shared memory");
    HSB_REQUIRE(groups <= 65535, HSB_ERR_UNSUPPORTED, "patch_conv_bwd_weight: groups > 65535");
    dim3 grid(B * fh * fw, groups);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
    if (dtype == HSB_F32) {
        e = cudaFuncSetAttribute(patch_conv_bwd_weight_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(HSB_ERR_CUDA, std::string("patch_conv_bwd_weight attr: ") + cudaGetErrorString(e));
        patch_conv_bwd_weight_kernel<float><<<grid, 256, smem, st>>>(p);
    } else {
        e = cudaFuncSetAttribute(patch_conv_bwd_weight_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(HSB_ERR_CUDA, std::string("patch_conv_bwd_weight attr: ") + cudaGetErrorString(e));
        patch_conv_bwd_weight_kernel<__nv_bfloat16><<<grid, 256, smem, st>>>(p);
    }
    return check_launch("patch_conv_bwd_weight launch");
}

extern "C" int hsb_patch_conv_bwd_input(const void* w, const void* dy, void* dx,
                                        int B, int Cin, int Cout, int H, int W, int fh, int fw,
                                        int kh, int kw, int pad_h, int pad_w, int dil_h, int dil_w, int groups,
                                        int pad_mode, int dtype, int w_layout, int64_t w_row_stride, void* stream) {
    HSB_REQUIRE(w && dy && dx, HSB_ERR_INVALID_ARG, "patch_conv_bwd_input: null pointer");
    PatchBwdParams p{};
    int rc = fill_patch_bwd(p, B, Cin, Cout, H, W, fh, fw, kh, kw, pad_h, pad_w, dil_h, dil_w, groups, pad_mode, dtype,
                            w_layout, w_row_stride, "patch_conv_bwd_input");
    if (rc != HSB_OK) return rc;
    p.w = w; p.dy = dy; p.dx = dx;
    const size_t total = (size_t)B * Cin * H * W;
    const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)std::max(1, device_sm_count()) * 32);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == HSB_F32) patch_conv_bwd_input_kernel<float><<<blocks, 256, 0, st>>>(p);
    else patch_conv_bwd_input_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(p);
    return check_launch("patch_conv_bwd_input launch");
}

static int fill_head_bwd(HeadBwdParams& p, int B, int sig_total, int sig_index, int sig_ch, int out_ch, int hp, int groups,
                         int fh, int fw, int64_t ssb, int64_t ssc, int64_t ssp, int dtype, int g_layout, int64_t g_row_stride,
                         const char* who) {
    HSB_REQUIRE(B > 0 && sig_total > 0 && sig_ch > 0 && out_ch > 0 && hp > 0 && groups > 0 && fh > 0 && fw > 0 && sig_index >= 0,
                HSB_ERR_INVALID_ARG, std::string(who) + ": bad dimension");
    HSB_REQUIRE(sig_ch % groups == 0 && out_ch % groups == 0 && hp <= out_ch && sig_index + sig_ch <= sig_total,
                HSB_ERR_INVALID_ARG, std::string(who) + ": bad channel split");
    HSB_REQUIRE(dtype == HSB_F32 || dtype == HSB_BF16, HSB_ERR_INVALID_ARG, std::string(who) + ": bad dtype");
    p.B = B; p.P = fh * fw; p.sig_total = sig_total; p.sig_index = sig_index; p.sig_ch = sig_ch; p.out_ch = out_ch; p.hp = hp;
    p.groups = groups; p.spg = sig_ch / groups; p.opg = out_ch / groups;
    p.ssb = ssb; p.ssc = ssc; p.ssp = ssp;
    if (g_layout == HSB_W_PATCH_MAJOR) {
        HSB_REQUIRE(g_row_stride >= hp, HSB_ERR_INVALID_ARG, std::string(who) + ": row stride < hp");
        p.gsb = (int64_t)p.P * g_row_stride; p.gsp = g_row_stride; p.gsk = 1;
    } else {
        p.gsb = (int64_t)hp * p.P; p.gsp = 1; p.gsk = p.P;
    }
    return HSB_OK;
}

extern "C" int hsb_signal2weights_bwd_signal(const void* ws, const void* dwout, void* ds,
                                             int B, int sig_total, int sig_index, int sig_ch, int out_ch, int hp, int groups,
                                             int fh, int fw, int64_t ds_stride_b, int64_t ds_stride_c, int64_t ds_stride_p,
                                             int dtype, int g_layout, int64_t g_row_stride, void* stream) {
    HSB_REQUIRE(ws && dwout && ds, HSB_ERR_INVALID_ARG, "signal2weights_bwd_signal: null pointer");
    HeadBwdParams p{};
    int rc = fill_head_bwd(p, B, sig_total, sig_index, sig_ch, out_ch, hp, groups, fh, fw, ds_stride_b, ds_stride_c, ds_stride_p,
                           dtype, g_layout, g_row_stride, "signal2weights_bwd_signal");
    if (rc != HSB_OK) return rc;
    p.ws = ws; p.dwout = dwout; p.ds = ds;
    const size_t total = (size_t)B * sig_total * p.P;
    const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)std::max(1, device_sm_count()) * 32);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == HSB_F32) head_bwd_signal_kernel<float><<<blocks, 256, 0, st>>>(p);
    else head_bwd_signal_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(p);
    return check_launch("signal2weights_bwd_signal launch");
}

extern "C" int hsb_signal2weights_bwd_weight(const void* s, const void* dwout, void* dws,
                                             int B, int sig_total, int sig_index, int sig_ch, int out_ch, int hp, int groups,
                                             int fh, int fw, int64_t s_stride_b, int64_t s_stride_c, int64_t s_stride_p,
                                             int dtype, int g_layout, int64_t g_row_stride, void* stream) {
    HSB_REQUIRE(s && dwout && dws, HSB_ERR_INVALID_ARG, "signal2weights_bwd_weight: null pointer");
    HeadBwdParams p{};
    int rc = fill_head_bwd(p, B, sig_total, sig_index, sig_ch, out_ch, hp, groups, fh, fw, s_stride_b, s_stride_c, s_stride_p,
                           dtype, g_layout, g_row_stride, "signal2weights_bwd_weight");
    if (rc != HSB_OK) return rc;
    p.s = s; p.dwout = dwout; p.dws = dws;
    const size_t warps = (size_t)out_ch * p.spg;
    const int blocks = (int)std::min<size_t>((warps * 32 + 255) / 256, (size_t)std::max(1, device_sm_count()) * 32);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == HSB_F32) head_bwd_weight_kernel<float><<<blocks, 256, 0, st>>>(p);
    else head_bwd_weight_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(p);
    return check_launch("signal2weights_bwd_weight launch");
}
