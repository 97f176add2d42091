"""Synthetic inputs for parity tests and benchmarks (numpy only, no torch, no reference).

Everything here is deterministic from plain numpy ``RandomState`` seeds, so the same rays and the
same network parameters can be rebuilt on any machine (the GPU box has no /root/reference and no
dataset).  Shapes follow SURVEY.md section 8(d):

* "lego-shaped" rays: H=W=800, camera_angle_x=0.6911112070083618, focal = .5*W/tan(.5*angle)
  (reference load_blender.py:99-100), pose = pose_spherical(theta, -30, 4.0)
  (load_blender.py:29-34,102), rays as in get_rays (run_nerf_helpers.py:162-171).
* NeRF parameters: same names/shapes as the reference ``NeRF`` module
  (run_nerf_helpers.py:76-103), values U(+-1/sqrt(fan_in)) like ``nn.Linear``'s default scale.
"""
import numpy as np

LEGO_CAMERA_ANGLE_X = 0.6911112070083618


def pose_spherical(theta, phi, radius):
    """c2w [4,4] float32; restates load_blender.py:9-34 (trans_t, rot_phi, rot_theta)."""
    t = np.eye(4, dtype=np.float32)
    t[2, 3] = radius
    p = phi / 180.0 * np.pi
    rp = np.array([[1, 0, 0, 0],
                   [0, np.cos(p), -np.sin(p), 0],
                   [0, np.sin(p), np.cos(p), 0],
                   [0, 0, 0, 1]], dtype=np.float32)
    th = theta / 180.0 * np.pi
    rt = np.array([[np.cos(th), 0, -np.sin(th), 0],
                   [0, 1, 0, 0],
                   [np.sin(th), 0, np.cos(th), 0],
                   [0, 0, 0, 1]], dtype=np.float32)
    flip = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.float32)
    return (flip @ (rt @ (rp @ t))).astype(np.float32)


def intrinsics(H, W, focal):
    """K as built in run_plnerf.py:1135-1140."""
    return np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]], dtype=np.float32)


def get_rays_np(H, W, K, c2w):
    """Restates get_rays_np (run_nerf_helpers.py:174-181) in float32."""
    i, j = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing="xy")
    dirs = np.stack([(i - K[0][2]) / K[0][0], -(j - K[1][2]) / K[1][1], -np.ones_like(i)], -1)
    rays_d = np.sum(dirs[..., np.newaxis, :] * c2w[:3, :3], -1).astype(np.float32)
    rays_o = np.broadcast_to(c2w[:3, -1], rays_d.shape).astype(np.float32)
    return rays_o, rays_d


def lego_rays(n_rays=None, H=800, W=800, theta=-180.0, seed=0):
    """Lego-shaped rays (rays_o, rays_d) float32 [n,3].  n_rays=None -> the full image in
    row-major pixel order; otherwise a seeded random subset of pixels (without replacement)."""
    focal = 0.5 * W / np.tan(0.5 * LEGO_CAMERA_ANGLE_X)
    K = intrinsics(H, W, focal)
    c2w = pose_spherical(theta, -30.0, 4.0)[:3, :4]
    ro, rd = get_rays_np(H, W, K, c2w)
    ro = ro.reshape(-1, 3)
    rd = rd.reshape(-1, 3)
    if n_rays is not None:
        idx = np.random.RandomState(seed).choice(H * W, size=n_rays, replace=False)
        ro, rd = ro[idx], rd[idx]
    return np.ascontiguousarray(ro), np.ascontiguousarray(rd), K, (H, W, focal)


def llff_rays(n_rays, H=378, W=504, focal=407.6, seed=0):
    """Fern-shaped forward-facing rays before NDC."""
    K = intrinsics(H, W, focal)
    c2w = np.eye(4, dtype=np.float32)[:3, :4].copy()
    c2w[:, 3] = [0.05, -0.02, 0.1]
    ro, rd = get_rays_np(H, W, K, c2w)
    idx = np.random.RandomState(seed).choice(H * W, size=n_rays, replace=False)
    return np.ascontiguousarray(ro.reshape(-1, 3)[idx]), np.ascontiguousarray(rd.reshape(-1, 3)[idx]), K, (H, W, focal)


def nerf_param_shapes(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=4, skips=(4,),
                      use_viewdirs=True):
    """Ordered (name, shape) list = ``NeRF.state_dict()`` of the reference
    (run_nerf_helpers.py:85-103)."""
    out = []
    for i in range(D):
        if i == 0:
            fan_in = input_ch
        elif (i - 1) in skips:
            fan_in = W + input_ch
        else:
            fan_in = W
        out.append((f"pts_linears.{i}.weight", (W, fan_in)))
        out.append((f"pts_linears.{i}.bias", (W,)))
    out.append(("views_linears.0.weight", (W // 2, input_ch_views + W)))
    out.append(("views_linears.0.bias", (W // 2,)))
    if use_viewdirs:
        out.append(("feature_linear.weight", (W, W)))
        out.append(("feature_linear.bias", (W,)))
        out.append(("alpha_linear.weight", (1, W)))
        out.append(("alpha_linear.bias", (1,)))
        out.append(("rgb_linear.weight", (3, W // 2)))
        out.append(("rgb_linear.bias", (3,)))
    else:
        out.append(("output_linear.weight", (output_ch, W)))
        out.append(("output_linear.bias", (output_ch,)))
    return out


def nerf_params(seed, density_boost=True, **kw):
    """Seeded float32 parameter dict.  ``density_boost`` applies SURVEY.md 8d's variant
    (sigma head weight*20, bias 0.5) so that transmittance actually decays along the ray."""
    rs = np.random.RandomState(seed)
    params = {}
    for name, shape in nerf_param_shapes(**kw):
        fan_in = shape[1] if len(shape) == 2 else None
        if fan_in is None:
            fan_in = params[name.replace("bias", "weight")].shape[1]
        bound = 1.0 / np.sqrt(fan_in)
        params[name] = rs.uniform(-bound, bound, size=shape).astype(np.float32)
    if density_boost:
        if kw.get("use_viewdirs", True):
            params["alpha_linear.weight"] *= 20.0
            params["alpha_linear.bias"][:] = 0.5
        else:
            params["output_linear.weight"][3] *= 20.0
            params["output_linear.bias"][3] = 0.5
    return params
