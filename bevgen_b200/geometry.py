"""Host-side (init-time) geometry for the stage-2 transformer: decode order, attention mask,
camera-bias prior and block layout.  Runs once per config on the CPU, exactly like the reference
(`GPTConfig.__post_init__`, multi_view_generation/modules/transformer/mingpt_sparse.py:74-102), but
written as closed forms over numpy arrays instead of the reference's scatter/gather construction.

Reference behaviour mirrored (file:line under /root/reference/multi_view_generation):
  decode order            modules/transformer/permuter.py:33-88   (CustomPermuter)
  allowed / window masks  modules/transformer/mask_generator.py:130-148,197-206
  camera-bias prior       mask_generator.py:150-190 (+ get_bev_weights :73-86, permuter.get_col_angles :153-162)
  block layout            mask_generator.py:192-228 (+ permuter.pattern_to_layout :98-123)
Quirks kept on purpose: swapped (img_h,img_w) in the ray helper call, the ray
x-component used as an "angle", rad2deg applied to a cosine *distance*.
"""
from enum import Enum

import numpy as np
import torch


class Cameras(Enum):
    """Camera name tuples; members and order as bev_utils/util.py:20-26."""
    NUSCENES_FRONT = ("CAM_FRONT",)
    NUSCENES_CAMERAS = ("CAM_FRONT", "CAM_BACK", "CAM_FRONT_RIGHT", "CAM_FRONT_LEFT", "CAM_BACK_RIGHT", "CAM_BACK_LEFT")
    NUSCENES_ABLATION_CAMERAS = ("CAM_FRONT", "CAM_FRONT_RIGHT", "CAM_FRONT_LEFT")
    ARGOVERSE_CAMERAS = ("ring_side_left", "ring_front_left", "ring_front_right", "ring_side_right")
    ARGOVERSE_FRONT_CAMERAS = ("ring_front_left", "ring_front_center", "ring_front_right")
    ARGOVERSE_ALL_CAMERAS = ("ring_side_left", "ring_front_left", "ring_front_center", "ring_front_right", "ring_side_right")

    def __getitem__(self, index):
        return self._value_[index]

    def __len__(self):
        return len(self._value_)

    def index(self, name):
        return self._value_.index(name)


class Dataset(Enum):
    NUSCENES = 0
    ARGOVERSE = 1


# (fx, fy, yaw) per nuScenes camera — calibration constants from permuter.py:151.
NUSCENES_CAM_DATA = {
    "CAM_FRONT": (1266.417203046554, 1266.417203046554, 0.005684811144346602),
    "CAM_BACK": (809.2209905677063, 809.2209905677063, 3.1391709219861887),
    "CAM_FRONT_RIGHT": (1260.8474446004698, 1260.8474446004698, 5.298742851167251),
    "CAM_FRONT_LEFT": (1272.5979470598488, 1272.5979470598488, 0.9627404474321728),
    "CAM_BACK_RIGHT": (1259.5137405846733, 1259.5137405846733, 4.349372983905386),
    "CAM_BACK_LEFT": (1256.7414812095406, 1256.7414812095406, 1.895431863668132),
}


def decode_order(num_cams, lat_h, lat_w, dataset, causal_order):
    """forward_shuffle_idx: token index in (cam,h,w) order of the t-th decoded token.

    nuScenes: per latent row, the front rig (FL, F, FR) and the back rig (BR, B, BL) are emitted
    centre-outward (centre camera split at its middle column, then the side cameras), left/right
    alternating, and the two rigs are interleaved element by element (permuter.py:48-69).
    Other datasets: row-major over cameras within each latent row (permuter.py:70-75).
    """
    n = num_cams * lat_h * lat_w
    if not causal_order:
        return np.arange(n, dtype=np.int64)

    def tok(cam, row, cols):
        return cam * lat_h * lat_w + row * lat_w + np.asarray(cols, dtype=np.int64)

    order = []
    if dataset == Dataset.NUSCENES:
        if num_cams == 3:
            rigs, names = [("CAM_FRONT_LEFT", "CAM_FRONT", "CAM_FRONT_RIGHT")], Cameras.NUSCENES_ABLATION_CAMERAS
        else:
            rigs = [("CAM_FRONT_LEFT", "CAM_FRONT", "CAM_FRONT_RIGHT"), ("CAM_BACK_RIGHT", "CAM_BACK", "CAM_BACK_LEFT")]
            names = Cameras.NUSCENES_CAMERAS
        mid = lat_w // 2
        cols = np.arange(lat_w)
        for row in range(lat_h):
            per_rig = []
            for left, centre, right in rigs:
                lc, cc, rc = names.index(left), names.index(centre), names.index(right)
                head = [] if lat_w % 2 == 0 else [tok(cc, row, [mid])]
                rstart = mid if lat_w % 2 == 0 else mid + 1
                go_left = np.concatenate([tok(cc, row, cols[:mid][::-1]), tok(lc, row, cols[::-1])])
                go_right = np.concatenate([tok(cc, row, cols[rstart:]), tok(rc, row, cols)])
                m = min(len(go_left), len(go_right))
                inter = np.stack([go_left[:m], go_right[:m]], 1).reshape(-1)
                per_rig.append(np.concatenate(head + [inter]))
            m = min(len(r) for r in per_rig)
            order.append(np.stack([r[:m] for r in per_rig], 1).reshape(-1))
    else:
        for row in range(lat_h):
            for cam in range(num_cams):
                order.append(tok(cam, row, np.arange(lat_w)))
    return np.concatenate(order).astype(np.int64)


def _pixel_ray_x(u, v, fx, fy, img_w, img_h):
    """x component of the unit ray through pixel (u,v) (bev_utils/nuscenes_helper.py:222-262)."""
    if not np.isclose(fx, fy, atol=5):
        raise ValueError(f"Focal lengths in the x and y directions must match: {fx} != {fy}")
    ray = np.array([u - img_w / 2, v - img_h / 2, fx], dtype=np.float64)
    return (ray / np.linalg.norm(ray))[0]


def column_angles(lat_w):
    """(6, lat_w) float32 table; the reference passes (img_h, img_w) swapped into the ray helper
    (permuter.py:158), reproduced here by swapping them in the call."""
    img_w, img_h = 1600, 900
    out = np.zeros((len(Cameras.NUSCENES_CAMERAS), lat_w), dtype=np.float32)
    for ci, name in enumerate(Cameras.NUSCENES_CAMERAS.value):
        fx, fy, yaw = NUSCENES_CAM_DATA[name]
        for c in range(lat_w):
            x = -_pixel_ray_x(img_w * ((c + 0.5) / lat_w), img_h / 2, fx, fy, img_h, img_w)
            out[ci, c] = np.float32(np.mod(yaw + x, 2 * np.pi))
    return out


def _cosine_cdist(a, b):
    """scipy.spatial.distance.cdist(a, b, 'cosine') in float64."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    na = np.sqrt((a * a).sum(1))[:, None]
    nb = np.sqrt((b * b).sum(1))[None, :]
    return 1.0 - (a @ b.T) / (na * nb)


def image_masks(num_img, window_len, fwd, causal_order):
    """(allowed, window) boolean (num_img, num_img) in *sequence* order (mask_generator.py:131-148).

    With causal_order the sequence order IS the decode order, so both are plain lower-triangular
    bands; otherwise position p(t) = argsort(fwd)[t] is the decode rank of token t.
    """
    pos = np.arange(num_img) if causal_order else np.argsort(fwd)
    pr, pc = pos[:, None], pos[None, :]
    allowed = pc <= pr
    window = allowed & (pc >= np.maximum(pr - window_len, 0))
    return allowed, window


def with_cond(img_block, num_cond, num_pad, cond_col_value, dtype):
    """Pad an (img,img) block to the full (L,L) sequence [cond | img | pad]: pad rows/cols 0, cond rows see
    no image columns, every row's cond columns = cond_col_value (mask_generator.pad_with_conf :68-71)."""
    n = img_block.shape[0]
    L = num_cond + n + num_pad
    out = np.zeros((L, L), dtype=dtype)
    out[num_cond:num_cond + n, num_cond:num_cond + n] = img_block
    out[:, :num_cond] = cond_col_value
    return out


def attention_mask(num_img, num_cond, num_pad, window_len, fwd, causal_order):
    allowed, _ = image_masks(num_img, window_len, fwd, causal_order)
    m = with_cond(allowed, num_cond, num_pad, True, bool)
    if num_pad:
        m[-num_pad:, 1:] = False          # keep exactly one unmasked column on pad rows (:203-205)
    return m


def bev_bearing_similarity(angles_decode_order, bev_h, bev_w):
    """(num_img, num_cond) float64 in [0,1]: cosine similarity between each image token's column
    bearing and each BEV cell's bearing around the ego centre (mask_generator.py:73-86)."""
    hh, ww = np.meshgrid(np.arange(bev_h), np.arange(bev_w), indexing="ij")
    # (n,2) interleaved float32 buffer with strided column views: torch's atan2 takes its scalar path on
    # strided operands and its SIMD path on contiguous ones, and the two differ in the last ulp; the
    # reference operates on column slices (mask_generator.py:78-83), so do the same to stay bit-identical.
    yx = torch.from_numpy(np.stack([hh.reshape(-1), ww.reshape(-1)], 1).astype(np.float32))
    yx[:, 0] *= -1
    yx[:, 0] += (bev_h // 2) - 0.5
    yx[:, 1] -= (bev_w // 2) - 0.5
    bev = torch.remainder(torch.atan2(yx[:, 0], yx[:, 1]) - torch.pi / 2, 2 * torch.pi).numpy()
    a = np.asarray(angles_decode_order)
    sim = 1.0 - _cosine_cdist(np.stack([np.cos(a), np.sin(a)], 1), np.stack([np.cos(bev), np.sin(bev)], 1))
    return (sim + 1.0) / 2.0


def camera_bias_prior(num_cams, lat_h, lat_w, bev_h, bev_w, num_pad, window_len, fwd, causal_order):
    """prob_matrix (L,L) float64 (mask_generator.py:150-190, legacy_prob_matrix=True path)."""
    num_img, num_cond = num_cams * lat_h * lat_w, bev_h * bev_w
    cam, row, col = np.meshgrid(np.arange(num_cams), np.arange(lat_h), np.arange(lat_w), indexing="ij")
    cam, row, col = cam.reshape(-1), row.reshape(-1), col.reshape(-1)
    ang = column_angles(lat_w)[cam, col]                                  # float32 per token, (cam,h,w) order
    unit = np.stack([np.cos(ang), np.sin(ang)], 1)                        # float32, as in the reference
    d_ang = np.rad2deg(_cosine_cdist(unit, unit))                          # the "BUG!!!" line (:156)
    d_row = np.abs(row[:, None].astype(np.float32) - row[None, :].astype(np.float32)).astype(np.float64)
    prob = np.exp(-0.5 * 4.0 ** (-2.0) * (d_ang + d_row))
    if causal_order:
        prob = prob[:, fwd][fwd, :]
    allowed, _ = image_masks(num_img, window_len, fwd, causal_order)
    prob[~allowed] = 0.0
    prob = np.clip(prob, 0.0, 1.0)
    full = with_cond(prob, num_cond, num_pad, 1.0, np.float64)
    sim = bev_bearing_similarity(ang[fwd], bev_h, bev_w)
    full[num_cond:num_cond + num_img, :num_cond] = sim
    return full


def layout_components(block, num_img, num_cond, num_pad, window_len, fwd, causal_order, prob_img):
    """The intermediates of mask_generator.outward_pattern (:192-214) that the per-head layouts are drawn from:
    (static block layout (nb, nb) bool = maxpool_block(window U cond columns), block prior (nb, nb) float32 =
    avgpool_block(padded prior), padded prior (L, L) float64 with 0.5 on the cond columns).
    `prob_img` is the (img,img) prior *before* cond padding (already permuted/masked/clamped)."""
    L = num_cond + num_img + num_pad
    nb = L // block
    _, window = image_masks(num_img, window_len, fwd, causal_order)
    static = with_cond(window, num_cond, num_pad, False, bool)
    if num_pad:
        static[-num_pad:, 0] = True
        static[-num_pad:, 1:] = False
    static_l = static.reshape(nb, block, nb, block).any(axis=(1, 3))
    pfull = with_cond(prob_img, num_cond, num_pad, 0.5, np.float64)
    prob_l = torch.nn.functional.avg_pool2d(torch.from_numpy(pfull)[None].to(torch.float), block, block)[0]
    return static_l, prob_l, pfull


def block_layouts(num_heads, block, density, num_img, num_cond, num_pad, window_len, fwd, causal_order, prob_img):
    """Per-head block layout (heads, L/block, L/block) int64 (mask_generator.py:192-228).

    static = maxpool_block(window ∪ cond columns); sampled = multinomial(avgpool_block(prob with 0.5 on
    cond columns), n) without replacement, restricted to blocks with non-zero prior.  At density 1.0
    every block with non-zero prior is drawn, so the result does not depend on the RNG.
    """
    static_l, prob_l, _ = layout_components(block, num_img, num_cond, num_pad, window_len, fwd, causal_order, prob_img)
    nb = static_l.shape[0]
    n_draw = int((nb * nb) * density - int(static_l.sum()))
    nonzero = (prob_l > 0).numpy()
    heads = []
    for _ in range(num_heads):
        sampled = np.zeros(nb * nb, dtype=bool)
        if n_draw > 0:
            if density >= 1.0 and n_draw >= int(nonzero.sum()):
                sampled = nonzero.reshape(-1).copy()
            else:
                idx = torch.multinomial(prob_l.flatten(), n_draw, replacement=False)   # same call as permuter.py:141
                sampled[idx.numpy()] = True
        sampled = sampled.reshape(nb, nb) & nonzero
        heads.append(static_l | sampled)
    return torch.from_numpy(np.stack(heads).astype(np.int64))
