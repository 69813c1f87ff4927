"""Host simulation of the HSB_IR_X5D operand assembly (patch_ir_tc.cu): emulates the three TMA boxes (zero fill outside
the image), the reflected rows, the halo-column copy and the constant-one channel exactly as the kernel indexes them,
then reads the buffer back as the MN-major A operand and compares it with the reflect-padded halo tile of the oracle.
Catches indexing mistakes without a GPU:   python scripts/probe/x5d_layout_sim.py"""
import numpy as np

rng = np.random.default_rng(0)


def simulate(CIN, PS, H, W, pi, pj, x):
    PH = PW = PS; TH = TW = PS + 2
    K1 = (CIN + 1 + 15) // 16 * 16
    XCH, BODY, BODY_UNITS, HALO = PW // 8, TH * PW, TH * (PW // 8), 2 * TH
    HALO_UNITS = (HALO + 7) // 8
    fh, fw = H // PH, W // PW

    def tma_box(chunk0, nchunks, row0):                       # [row][chunk][channel K1][8 px], zero fill outside
        out = np.zeros((TH, nchunks, K1, 8), dtype=np.float32)
        for r in range(TH):
            for ch in range(nchunks):
                gy, gc = row0 + r, chunk0 + ch
                if 0 <= gy < H and 0 <= gc < W // 8:
                    out[r, ch, :CIN, :] = x[:, gy, gc * 8:gc * 8 + 8]
        return out

    units = np.full((BODY_UNITS + HALO_UNITS, K1, 8), 777.0, dtype=np.float32)      # stale bytes of the other tenant
    cy, cx = pi * PH - 1, pj * XCH
    units[:BODY_UNITS] = tma_box(cx, XCH, cy).reshape(BODY_UNITS, K1, 8)
    hl, hr = tma_box(cx - 1, 1, cy)[:, 0], tma_box(cx + XCH, 1, cy)[:, 0]          # [row][K1][8]
    left, right, top, bottom = pj == 0, pj == fw - 1, pi == 0, pi == fh - 1
    a1 = units
    for ch in range(XCH):                                                           # reflected rows (channels < CIN only)
        for c in range(CIN):
            if top: a1[0 * XCH + ch, c] = a1[2 * XCH + ch, c].copy()
            if bottom: a1[(TH - 1) * XCH + ch, c] = a1[(TH - 3) * XCH + ch, c].copy()
    src_body = units.copy()      # the kernel reads rows rs that step 1 never writes, so a snapshot is equivalent
    for h in range(HALO):
        side = 1 if h >= TH else 0
        r = h - side * TH
        rs = 2 if (top and r == 0) else (TH - 3 if (bottom and r == TH - 1) else r)
        for c in range(CIN):
            if side == 0:
                v = src_body[rs * XCH, c, 1] if left else hl[rs, c, 7]
            else:
                v = src_body[rs * XCH + XCH - 1, c, 6] if right else hr[rs, c, 0]
            a1[BODY_UNITS + (h >> 3), c, h & 7] = v
    for u in range(BODY_UNITS + HALO_UNITS):
        a1[u, CIN, :] = 1.0
    # operand view: A[m][k] = unit m // 8, channel k, pixel m % 8
    T = TH * TW
    A = np.stack([a1[m // 8, :, m % 8] for m in range(T)])
    # expected: reflect-padded tile, M order = body (row, col 1..PW) then left column, then right column
    xp = np.pad(x, ((0, 0), (1, 1), (1, 1)), mode="reflect")
    tile = xp[:, pi * PH:pi * PH + TH, pj * PW:pj * PW + TW]                       # (C, TH, TW)
    exp = np.zeros((T, CIN + 1), dtype=np.float32)
    for m in range(T):
        if m < BODY: r, c = m // PW, m % PW + 1
        else:
            h = m - BODY; side = 1 if h >= TH else 0; r, c = h - side * TH, (TW - 1 if side else 0)
        exp[m, :CIN] = tile[:, r, c]; exp[m, CIN] = 1.0
    assert np.array_equal(A[:, :CIN + 1], exp), (CIN, PS, pi, pj)
    # hidden-tile position of M row m (epilogue 1) must enumerate the tile exactly once
    pos = []
    for m in range(T):
        if m < BODY: pos.append((m // PW) * TW + m % PW + 1)
        else:
            h = m - BODY; side = 1 if h >= TH else 0; pos.append((h - side * TH) * TW + (TW - 1 if side else 0))
    assert sorted(pos) == list(range(T))


for CIN, PS in [(34, 16), (26, 16), (22, 16), (24, 8), (14, 8)]:
    for (fh, fw) in [(3, 3), (1, 1), (2, 1), (1, 3)]:
        H, W = fh * PS, fw * PS
        x = rng.standard_normal((CIN, H, W)).astype(np.float32)
        for pi in range(fh):
            for pj in range(fw):
                simulate(CIN, PS, H, W, pi, pj, x)
print("x5d operand assembly matches the reflect-padded halo tile for every patch position and shape")
