"""Host statement of the persistent decode kernel's weight format (bevgen_b200/decode_format.py): round trip and the lane / register
mapping the kernel's mma.m16n8k16 consumption relies on, checked against a plain matmul (no GPU)."""
import numpy as np

from bevgen_b200 import decode_format as F


def test_pack_unpack_round_trip_and_three_byte_accuracy():
    rng = np.random.default_rng(0)
    for rows, d, nq in ((24, 64, 1), (16, 128, 4), (13, 64, 1)):
        w = (rng.standard_normal((rows, nq * d)) * 0.02).astype(np.float32)
        amax = float(np.abs(w).max())
        lo_mul = 2.0 ** (19 - int(np.floor(np.log2(amax))))
        packed = F.pack_reference(w, d, nq, lo_mul)
        assert packed.size == nq * ((rows + 7) // 8) * (d // 64) * F.KG_BYTES
        w16, lo = F.unpack(packed, rows, d, nq, lo_mul)
        assert np.array_equal(w16, w.astype(np.float16).astype(np.float32))
        rel = np.abs((w16 + lo) - w).max() / amax
        assert rel < 2.0 ** -14, rel                 # fp16 + e4m3 residual: ~2^-15 of the largest weight
        assert np.abs(w16 - w).max() / amax > 2.0 ** -13       # (the fp16 plane alone is 8x worse)


def test_emulated_unit_mma_matches_matmul():
    rng = np.random.default_rng(1)
    d = 128
    w = (rng.standard_normal((8, d)) * 0.02).astype(np.float32)
    x = rng.standard_normal((16, d)).astype(np.float32)
    lo_mul = 2.0 ** (19 - int(np.floor(np.log2(float(np.abs(w).max())))))
    packed = F.pack_reference(w, d, 1, lo_mul)
    got = F.emulate_unit_mma(packed, x, d, lo_mul)
    want = x.astype(np.float64) @ w.astype(np.float64).T
    assert np.abs(got - want).max() < 3e-5 * np.abs(want).max() + 1e-6
