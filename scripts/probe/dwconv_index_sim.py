"""Host emulation of dwconv_bias_act_nhwc_kernel's indexing (csrc/encoder_epilogue.cu): strips of DW_P output pixels,
tap / padding arithmetic, chunk assignment -- against F.pad + F.conv2d on the CPU.   python scripts/probe/dwconv_index_sim.py"""
import numpy as np
import torch
import torch.nn.functional as F

DW_P, MAX_CHUNKS = 2, 64


def emulate(x, w, bias, k, S, pad_t, pad_l, Ho, Wo):
    N, C, H, W = x.shape
    G = C // 8
    L = 1 if G >= 256 else 256 // G
    strips_x = (Wo + DW_P - 1) // DW_P
    strips = Ho * strips_x
    chunks = min((strips + L - 1) // L, MAX_CHUNKS)
    y = np.zeros((N, C, Ho, Wo), dtype=np.float64)
    pool = np.zeros((N, chunks, C))
    IN = (DW_P - 1) * S + k
    for n in range(N):
        for bx in range(chunks):
            for l in range(L):
                s = bx * L + l
                while s < strips:
                    oy, ox0 = s // strips_x, (s % strips_x) * DW_P
                    acc = np.tile(bias[None, :], (DW_P, 1)).astype(np.float64)
                    for ky in range(k):
                        iy = oy * S - pad_t + ky
                        if iy < 0 or iy >= H:
                            continue
                        inp = np.zeros((IN, C))
                        for j in range(IN):
                            ix = ox0 * S - pad_l + j
                            if 0 <= ix < W:
                                inp[j] = x[n, :, iy, ix]
                        for kx in range(k):
                            for q in range(DW_P):
                                acc[q] += w[:, ky, kx] * inp[q * S + kx]
                    for q in range(DW_P):
                        if ox0 + q < Wo:
                            y[n, :, oy, ox0 + q] = acc[q]
                            pool[n, bx] += acc[q]
                    s += chunks * L
    return y, pool


rng = np.random.default_rng(0)
for (N, C, H, W, k, S, pads) in [(1, 16, 9, 11, 3, 2, (0, 1, 0, 1)), (2, 24, 7, 10, 3, 1, (1, 1, 1, 1)), (1, 16, 8, 9, 5, 2, (1, 2, 1, 2)),
                                 (1, 2048 // 8, 5, 6, 5, 1, (2, 2, 2, 2)), (1, 16, 70, 40, 3, 1, (1, 1, 1, 1))]:
    x = rng.standard_normal((N, C, H, W)); w = rng.standard_normal((C, k, k)); b = rng.standard_normal(C)
    ref = F.conv2d(F.pad(torch.from_numpy(x), pads), torch.from_numpy(w)[:, None], torch.from_numpy(b), stride=S, groups=C).numpy()
    Ho, Wo = ref.shape[-2:]
    y, pool = emulate(x, w, b, k, S, pads[2], pads[0], Ho, Wo)
    assert np.allclose(y, ref, atol=1e-10), (C, k, S)
    assert np.allclose(pool.sum(1), ref.sum((2, 3)), atol=1e-8)
print("fused depthwise indexing matches F.conv2d for every geometry (incl. asymmetric SAME padding, ragged strips, > 64 chunks)")
