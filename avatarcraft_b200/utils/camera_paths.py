"""Camera orbit helpers of the entry points (reference: utils/render_utils.py:137-154 default_360_path,
:323-337 pose2cap, :363-376 cap2rays; render_canonical.py:60-71 intrinsics f = 0.78125 W, c = W/2)."""
import numpy as np

from .synthetic import orbit_pose, pinhole_rays


def default_360_path(center, dist, trajectory_resolution=60):
    """`trajectory_resolution` camera-to-world poses on a horizontal circle of radius `dist` around `center`."""
    return [orbit_pose(a, dist=dist, center=center) for a in np.linspace(-180.0, 180.0, trajectory_resolution, endpoint=False)]


def rays_for_pose(c2w, width, height, device):
    o, d = pinhole_rays(c2w, width, height)
    return o.to(device), d.to(device)
