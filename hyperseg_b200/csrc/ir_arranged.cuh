// "Arranged" per-patch weight rows of the fused inverted-residual MetaBlock, and the geometry of the kernel that
// consumes them (patch_ir2.cu).  Shared by that kernel, the raw -> arranged re-arrangement kernel and the weight head
// that emits arranged rows directly (signal2weights_arranged in signal2weights_tc.cu).
//
// The reference's weight vector of a patch is [W1 (hid x Cin) | W2 (hid x 3 x 3) | W3 (Cout x hid)]
// (hyperseg/models/hyperseg_v1_0.py:301-309, :350, :357-358, :364).  An arranged row holds the same numbers, with the
// eval-mode BatchNorm scales folded in, in the order the tensor core and the depthwise stage read them:
//
//   B1   [kc < KC1][n < hid][8]   bf16   = s1[n] * W1[n][8 kc + j]      (0 for 8 kc + j >= Cin)   K-major UMMA operand,
//                                                                        rows 16 B apart, 8-channel chunks hid*16 B apart
//   W2T  [tap < 9][c < hid]       bf16   = s2[c] * W2[c][tap]           (padded to a multiple of 16 bytes)
//   B2   [kc < KC2][n < Cout][8]  bf16   = s3[n] * W3[n][8 kc + j]      (0 for 8 kc + j >= hid)
//
// The BatchNorm shifts are not part of the row: they are static and enter through a constant "init" MMA
// (GEMM1 / GEMM2) or as the initial value of the depthwise accumulation.
#pragma once
#include <stdint.h>

namespace hsb {

constexpr int ir_r16(int v) { return (v + 15) / 16 * 16; }
constexpr int ir_r8(int v) { return (v + 7) / 8 * 8; }
constexpr int ir_r128(int v) { return (v + 127) / 128 * 128; }
constexpr int ir_r1024(int v) { return (v + 1023) / 1024 * 1024; }
constexpr int ir_max(int a, int b) { return a > b ? a : b; }
constexpr int ir_min(int a, int b) { return a < b ? a : b; }
constexpr int ir_pow2_cols(int v) { int p = 32; while (p < v) p *= 2; return p; }

// layout of one arranged row (bytes); usable from host code with run-time dimensions
struct IRRow {
    int kc1, kc2, b1_lbo, b2_lbo, sz_b1, sz_w2t, sz_b2, bytes;
    __host__ __device__ constexpr IRRow(int cin, int hid, int cout)
        : kc1((cin + 7) / 8), kc2((hid + 7) / 8), b1_lbo(hid * 16), b2_lbo(cout * 16),
          sz_b1(((cin + 7) / 8) * hid * 16), sz_w2t(ir_r16(18 * hid)), sz_b2(((hid + 7) / 8) * cout * 16),
          bytes(((cin + 7) / 8) * hid * 16 + ir_r16(18 * hid) + ((hid + 7) / 8) * cout * 16) {}
};

// Which (reference weight index, BatchNorm scale index) feeds arranged element e of a row: src < 0 means zero padding.
// which = 0/1/2 selects the scale vector (bn1 / bn2 / bn3).
struct IRSource { int src; int which; int ch; };
__host__ __device__ inline IRSource ir_arranged_source(int e, int cin, int hid, int cout) {
    const IRRow r(cin, hid, cout);
    const int e1 = r.sz_b1 / 2, e2 = e1 + r.sz_w2t / 2;
    if (e < e1) {
        const int kc = e / (hid * 8), rem = e % (hid * 8), n = rem / 8, k = kc * 8 + rem % 8;
        return {k < cin ? n * cin + k : -1, 0, n};
    }
    if (e < e2) {
        const int t = e - e1;
        if (t >= 9 * hid) return {-1, 1, 0};
        const int tap = t / hid, c = t % hid;
        return {cin * hid + c * 9 + tap, 1, c};
    }
    const int t = e - e2, kc = t / (cout * 8), rem = t % (cout * 8), n = rem / 8, k = kc * 8 + rem % 8;
    return {k < hid ? cin * hid + 9 * hid + n * hid + k : -1, 2, n};
}

}  // namespace hsb
