"""Host-side description of the persistent decode kernel's weight format (csrc/decode_persistent.cu, bevgen_pack_decode_linear).

A linear layer W [rows][K_total] is cut into UNITS of 8 weight rows x d columns (MLP2: K_total = 4 d -> 4 column quarters, unit index =
quarter * rows/8 + row_unit).  A unit is d/64 K-GROUPS of 1536 bytes; a k-group holds, for the 64 columns k0..k0+63 and the 32 lanes of
the consuming warp (lane = 4 g + t: g = weight row within the unit, t = column pair), the B fragments of four mma.m16n8k16 k-steps:

    bytes    0 ..  511   fp16, k-steps 0,1 : lane * 16 + {ks0: W[g][2t], W[g][2t+1], W[g][2t+8], W[g][2t+9]; ks1: same at +16}
    bytes  512 .. 1023   fp16, k-steps 2,3
    bytes 1024 .. 1535   e4m3 of (W - fp16(W)) * lo_mul, lane * 16 + 4 bytes per k-step in the same column order

so that every lane fetches its operands with three conflict-free 16-byte shared-memory loads and a CTA's units are one contiguous
byte range (one cp.async.bulk per unit).  This module is the numpy statement of that layout: `pack_reference` is what the CUDA pack
kernel must produce, `unpack` inverts it (used by the tests; nothing here runs on the hot path)."""
import numpy as np

KG_BYTES = 1536


def _cols(t):
    return [2 * t, 2 * t + 1, 2 * t + 8, 2 * t + 9]


def e4m3_encode(x):
    """float32 array -> uint8 e4m3fn (round to nearest even, saturating at +-448), via torch."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).clamp(-448.0, 448.0).to(torch.float8_e4m3fn)
    return t.view(torch.uint8).numpy()


def e4m3_decode(b):
    import torch
    return torch.from_numpy(np.ascontiguousarray(b, dtype=np.uint8)).view(torch.float8_e4m3fn).float().numpy()


def pack_reference(w, d, n_quarters=1, lo_mul=1.0):
    """w [rows][n_quarters * d] float32 -> uint8 [units * (d // 64) * 1536]."""
    rows, ld = w.shape
    assert ld >= n_quarters * d and d % 64 == 0
    upq = (rows + 7) // 8
    KG = d // 64
    wp = np.zeros((upq * 8, n_quarters * d), dtype=np.float32)
    wp[:rows] = w[:, : n_quarters * d]
    w16 = wp.astype(np.float16)
    res8 = e4m3_encode((wp - w16.astype(np.float32)) * np.float32(lo_mul))
    out = np.zeros((n_quarters * upq, KG, KG_BYTES), dtype=np.uint8)
    for q in range(n_quarters):
        for ru in range(upq):
            u = q * upq + ru
            for kg in range(KG):
                base = q * d + kg * 64
                for lane in range(32):
                    g, t = lane >> 2, lane & 3
                    r = ru * 8 + g
                    for ks in range(4):
                        cols = [base + ks * 16 + c for c in _cols(t)]
                        hi = w16[r, cols].view(np.uint8)                       # 8 bytes
                        part, e = ks >> 1, ks & 1
                        o = part * 512 + lane * 16 + e * 8
                        out[u, kg, o:o + 8] = hi
                        out[u, kg, 1024 + lane * 16 + ks * 4: 1024 + lane * 16 + ks * 4 + 4] = res8[r, cols]
    return out.reshape(-1)


def unpack(packed, rows, d, n_quarters=1, lo_mul=1.0):
    """Inverse of the packing: -> (w16 [rows][n_quarters*d] float32 values of the fp16 plane, residual plane / lo_mul)."""
    upq = (rows + 7) // 8
    KG = d // 64
    p = np.asarray(packed, dtype=np.uint8).reshape(n_quarters * upq, KG, KG_BYTES)
    w16 = np.zeros((upq * 8, n_quarters * d), dtype=np.float32)
    lo = np.zeros_like(w16)
    for q in range(n_quarters):
        for ru in range(upq):
            u = q * upq + ru
            for kg in range(KG):
                base = q * d + kg * 64
                for lane in range(32):
                    g, t = lane >> 2, lane & 3
                    r = ru * 8 + g
                    for ks in range(4):
                        cols = [base + ks * 16 + c for c in _cols(t)]
                        part, e = ks >> 1, ks & 1
                        o = part * 512 + lane * 16 + e * 8
                        w16[r, cols] = p[u, kg, o:o + 8].copy().view(np.float16).astype(np.float32)
                        lo[r, cols] = e4m3_decode(p[u, kg, 1024 + lane * 16 + ks * 4: 1024 + lane * 16 + ks * 4 + 4]) / np.float32(lo_mul)
    return w16[:rows], lo[:rows]


def emulate_unit_mma(packed_unit, x, d, lo_mul):
    """What one CTA computes for one unit: packed_unit uint8 [(d//64)*1536], x [16][d] float32 -> out [16][8], following the kernel's
    lane / register mapping (A fragment rows = batch, B fragment column = weight row g; accumulator c0..c3 = (g,2t),(g,2t+1),(g+8,2t),(g+8,2t+1))."""
    KG = d // 64
    p = np.asarray(packed_unit, dtype=np.uint8).reshape(KG, KG_BYTES)
    xhi = x.astype(np.float16)
    xlo = (x - xhi.astype(np.float32)).astype(np.float16)
    out = np.zeros((16, 8), dtype=np.float64)
    for w in range(KG):                      # warp w owns k-group w
        for ks in range(4):
            k0 = w * 64 + ks * 16
            # B[k][n]: k in 0..15, n in 0..7 assembled from the lanes' registers
            Bh = np.zeros((16, 8), dtype=np.float32)
            Bl = np.zeros((16, 8), dtype=np.float32)
            for lane in range(32):
                g, t = lane >> 2, lane & 3
                part, e = ks >> 1, ks & 1
                o = part * 512 + lane * 16 + e * 8
                h = p[w, o:o + 8].copy().view(np.float16).astype(np.float32)          # b0 = (k 2t, 2t+1), b1 = (k 2t+8, 2t+9), column n = g
                l = e4m3_decode(p[w, 1024 + lane * 16 + ks * 4: 1024 + lane * 16 + ks * 4 + 4])
                for i, k in enumerate(_cols(t)):
                    Bh[k, g] = h[i]
                    Bl[k, g] = l[i]
            A_hi = xhi[:, k0:k0 + 16].astype(np.float32)
            A_lo = xlo[:, k0:k0 + 16].astype(np.float32)
            out += (A_hi @ Bh).astype(np.float64) + (A_lo @ Bh).astype(np.float64) + (A_hi @ Bl).astype(np.float64) / lo_mul
    return out
