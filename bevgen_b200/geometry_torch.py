"""Small init-time tensors of the stage-2 embeddings (host side, computed once):
image-plane pixel grid (mingpt_sparse.py:256-264,288-292) and the BEV ego-frame grid (get_bev_grid :116-141)."""
import torch
import torch.nn.functional as F


def generate_grid(height: int, width: int):
    xs = torch.linspace(0, 1, width)
    ys = torch.linspace(0, 1, height)
    g = torch.stack(torch.meshgrid((xs, ys), indexing="xy"), 0)      # 2 h w
    return F.pad(g, (0, 0, 0, 0, 0, 1), value=1)[None]               # 1 3 h w


def image_plane(lat_h, lat_w, cam_res):
    """[hw, 3] rows (x*cam_res[0], y*cam_res[1], 1) — note the reference scales x by cam_res[0] and y by cam_res[1]."""
    g = generate_grid(lat_h, lat_w)[0].clone()
    g[0] *= cam_res[0]
    g[1] *= cam_res[1]
    return g.reshape(3, lat_h * lat_w).t().contiguous()


def bev_grid(h, w, offset=0):
    grid = generate_grid(h, w).squeeze(0)
    grid[0] = w * grid[0]
    grid[1] = h * grid[1]
    sh, sw = h / 80, w / 80
    V = torch.tensor([[0., -sw, w / 2.], [-sh, 0., h * offset + h / 2.], [0., 0., 1.]])
    return (V.inverse() @ grid.reshape(3, -1)).reshape(3, h, w)
