"""Host-side SMPL body model: per-vertex rest->pose 4x4 transforms for the warp path.

Restates what render_warp.py needs from the reference's models/smpl.py: `batch_rodrigues` (:549-580, with its
`+1e-8` inside the norm), `batch_rigid_transform` (:596-647), `lbs(..., return_T=True)` (:351-437 -- note that it
DROPS the pose blend-shapes, `v_posed = v_shaped`, :420) and `SMPL.verts_transformations` (:107-161).
This runs once per frame on the host in float32 torch, exactly like the reference (`device='cpu'`,
render_warp.py:133-138); the per-sample work it feeds is on the GPU (utils/ray_utils.py).

The licensed SMPL_NEUTRAL.pkl is not part of the repository (readme.md:41-47): the class takes either a path to
that pickle or a dict with the same fields (utils/synthetic.synthetic_smpl_model for tests)."""
import pickle

import numpy as np
import torch
import torch.nn.functional as F

SMPL_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21]


def batch_rodrigues(rot_vecs):
    """[N,3] axis-angle -> [N,3,3]."""
    angle = torch.norm(rot_vecs + 1e-8, dim=1, keepdim=True)
    axis = rot_vecs / angle
    c, s = torch.cos(angle)[:, None], torch.sin(angle)[:, None]
    rx, ry, rz = axis[:, 0:1], axis[:, 1:2], axis[:, 2:3]
    zero = torch.zeros_like(rx)
    K = torch.cat([zero, -rz, ry, rz, zero, -rx, -ry, rx, zero], dim=1).view(-1, 3, 3)
    return torch.eye(3, dtype=rot_vecs.dtype)[None] + s * K + (1 - c) * torch.bmm(K, K)


def batch_rigid_transform(rot_mats, joints, parents):
    """rot_mats [B,J,3,3], joints [B,J,3] -> posed joints [B,J,3], relative transforms A [B,J,4,4]."""
    B, J = joints.shape[:2]
    joints = joints.unsqueeze(-1)
    rel = joints.clone()
    rel[:, 1:] -= joints[:, parents[1:]]
    local = torch.cat([F.pad(rot_mats.reshape(-1, 3, 3), [0, 0, 0, 1]), F.pad(rel.reshape(-1, 3, 1), [0, 0, 0, 1], value=1)],
                      dim=2).view(B, J, 4, 4)
    chain = [local[:, 0]]
    for i in range(1, J):
        chain.append(torch.matmul(chain[int(parents[i])], local[:, i]))
    G = torch.stack(chain, dim=1)
    A = G - F.pad(torch.matmul(G, F.pad(joints, [0, 0, 0, 1])), [3, 0, 0, 0, 0, 0, 0, 0])
    return G[:, :, :3, 3], A


class SMPL:
    def __init__(self, model, device="cpu"):
        if isinstance(model, str):
            with open(model, "rb") as f:
                model = pickle.load(f, encoding="latin1")
        t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float32))
        self.v_template = t(model["v_template"])                                  # [V,3]
        self.shapedirs = t(np.asarray(model["shapedirs"])[:, :, :10])             # [V,3,10]
        jr = model["J_regressor"]
        self.J_regressor = t(jr.todense() if hasattr(jr, "todense") else jr)      # [24,V]
        parents = np.asarray(model["kintree_table"])[0].astype(np.int64).copy() if "kintree_table" in model else np.asarray(SMPL_PARENTS)
        parents[0] = -1
        self.parents = torch.as_tensor(parents)
        self.lbs_weights = t(model["weights"])                                    # [V,24]
        self.faces = np.asarray(model["f"], dtype=np.int32) if "f" in model else None

    def _lbs(self, betas, pose):
        betas, pose = torch.as_tensor(betas, dtype=torch.float32), torch.as_tensor(pose, dtype=torch.float32)
        v_delta = torch.einsum("bl,mkl->bmk", betas, self.shapedirs)
        v_shaped = self.v_template[None] + v_delta
        J = torch.einsum("bik,ji->bjk", v_shaped, self.J_regressor)
        rot = batch_rodrigues(pose.view(-1, 3)).view(1, -1, 3, 3)
        J_posed, A = batch_rigid_transform(rot, J, self.parents)
        T = torch.matmul(self.lbs_weights[None], A.view(1, -1, 16)).view(1, -1, 4, 4)
        return v_shaped, v_delta, J, J_posed, A, T

    def verts_transformations(self, poses, betas, return_tensor=True, concat_joints=False):
        """(v_shaped[+J], T[+A], v_delta): per-vertex blended joint transforms; the returned vertices are the
        SHAPED, UN-posed template (the reference's lbs(return_T=True) behaviour, models/smpl.py:420-433)."""
        assert np.asarray(poses).shape[0] == 1
        v_shaped, v_delta, J, _, A, T = self._lbs(betas, poses)
        verts = torch.cat([v_shaped, J], dim=1) if concat_joints else v_shaped
        Tall = torch.cat([T, A], dim=1) if concat_joints else T
        if not return_tensor:
            return verts.numpy()[0], Tall.numpy()[0], v_delta
        return verts, Tall, v_delta

    def forward(self, poses, betas, return_tensor=True, return_joints=False):
        """Posed vertices (and joints) by linear blend skinning."""
        v_shaped, _, _, J_posed, _, T = self._lbs(betas, poses)
        vh = F.pad(v_shaped, [0, 1], value=1.0)
        verts = torch.matmul(T, vh.unsqueeze(-1))[:, :, :3, 0]
        if not return_tensor:
            verts, J_posed = verts.numpy()[0], J_posed.numpy()[0]
        return (verts, J_posed) if return_joints else verts


def calc_local_trans(body_model, poses=None, shape_from=None, shape_to=None, render_type="animate", n_interp=10, max_frames=100,
                     scale=1.0, smpl_scale=0.9):
    """Per frame: posed SMPL surface and the per-vertex rest->scene transforms the warp inverts
    (render_warp.py:127-222).  Returns (world_verts [F][6890,3] f32, Ts [F][6914,4,4] f64, n_frames)."""
    zero_shape = np.zeros((1, 10), np.float32)
    if render_type == "animate":
        n_frame = min(max_frames, poses.shape[0])
        shapes = np.zeros((n_frame, 1, 10), np.float32)
    elif render_type == "interp_shape":
        shapes = np.linspace(shape_from, shape_to, n_interp).astype(np.float32)
        n_frame = min(max_frames, shapes.shape[0])
        poses = np.zeros((n_frame, 72), np.float32)
    else:
        raise NotImplementedError
    da = np.zeros((24, 3), np.float32)                       # the "da" (大) rest pose of NeuMan
    da[1], da[2] = [0, 0, 1.0], [0, 0, -1.0]
    da = da.reshape(1, 72)
    v0, T_t2rest, _ = body_model.verts_transformations(da, zero_shape, return_tensor=False, concat_joints=True)
    rest_verts, rest_joints = body_model.forward(da, zero_shape, return_tensor=False, return_joints=True)
    rest_h = np.concatenate([np.concatenate([rest_verts, rest_joints], 0), np.ones((rest_verts.shape[0] + 24, 1))], 1)
    world_verts, Ts = [], []
    for i in range(n_frame):
        _, T_t2pose, _ = body_model.verts_transformations(poses[i][None], zero_shape, return_tensor=False, concat_joints=True)
        vt, _, _ = body_model.verts_transformations(da, shapes[i].reshape(1, 10), return_tensor=False, concat_joints=True)
        T_shape = np.tile(np.eye(4), (v0.shape[0], 1, 1))
        T_shape[:, :3, 3] += (v0 - vt)
        T_rest2pose = T_t2pose @ np.linalg.inv(T_shape) @ np.linalg.inv(T_t2rest)
        S = np.eye(4)
        S[:3, :3] *= scale
        T_scene = S @ T_rest2pose
        Ts.append(T_rest2pose @ (np.eye(4) / smpl_scale))     # NB divides the homogeneous entry too (render_warp.py:200-204)
        world_verts.append(np.einsum("nij,nj->ni", T_scene, rest_h)[:6890, :3].astype(np.float32))
    return world_verts, Ts, n_frame
