"""CPU restatement of poseRefine::process (reference: linemodLevelup/linemodLevelup.cpp:27-155) and of
the Open3D calls it makes -- TEST INFRASTRUCTURE ONLY (see oracle/lm_oracle.cpp header for who may
import oracle/).

PARITY UNPINNED.  All ICP arithmetic of the reference lives in Open3D, which is neither vendored nor
pinned (find_package(Open3D REQUIRED), linemodLevelup/CMakeLists.txt:13; API shape => 0.8/0.9), is
absent from this container, and no reference test holds a poseRefine output.  The Open3D pieces are
restated from its published algorithm (open3d/geometry/PointCloud.cpp VoxelDownSample,
EstimateNormals.cpp ComputeNormal + KDTreeSearchParamKNN(30), registration/Registration.cpp
RegistrationICP, TransformationEstimation.cpp PointToPlane, utility/Eigen.cpp
SolveJacobianSystemAndObtainExtrinsicMatrix / TransformVector6dToMatrix4d).  Known freedoms that do
not affect the result beyond rounding: voxel output order (hash-map order there, sorted here),
normal sign, OpenMP reduction order.

The reference's bug at LL.cpp:109 is reproduced on purpose: the ICP TARGET is the down-sampled MODEL
cloud, the scene depth only enters through the initial translation (LL.cpp:101-104) and the bounds
check (LL.cpp:52-55).
"""
import numpy as np
from scipy.spatial import cKDTree

VOXEL = 0.0025      # LL.cpp:106
MAX_DIST = 0.01     # LL.cpp:31
KNN = 30            # Open3D EstimateNormals default
REL_FITNESS = 1e-6  # ICPConvergenceCriteria defaults
REL_RMSE = 1e-6
MAX_ITER = 30


def model_bbox(model_depth, half=4):
    """dilate(modelDepth > 0, 9x9 ones) then boundingRect(findNonZero): LL.cpp:43-50."""
    ys, xs = np.nonzero(model_depth > 0)
    if ys.size == 0:
        return 0, 0, 0, 0
    H, W = model_depth.shape
    x0, x1 = max(xs.min() - half, 0), min(xs.max() + half, W - 1)
    y0, y1 = max(ys.min() - half, 0), min(ys.max() + half, H - 1)
    return int(x0), int(y0), int(x1 - x0 + 1), int(y1 - y0 + 1)


def build_clouds(scene_depth, model_depth, sceneK, modelK, detectX, detectY, bbox, half=4):
    """LL.cpp:57-104.  Returns model points (N,3) f64 row-major order, center_model, center_scene."""
    bx, by, bw, bh = bbox
    sceneK = np.asarray(sceneK, np.float32)
    modelK = np.asarray(modelK, np.float32)
    rr, cc = np.mgrid[0:bh, 0:bw]
    mr, mc = rr + by, cc + bx
    sr = np.maximum(rr + detectY - half, 0)
    sc = np.maximum(cc + detectX - half, 0)
    md = model_depth[mr, mc]
    # inside the dilated mask
    pad = np.pad(model_depth > 0, half)
    dil = np.zeros(model_depth.shape, bool)
    for dy in range(2 * half + 1):
        for dx in range(2 * half + 1):
            dil |= pad[dy:dy + model_depth.shape[0], dx:dx + model_depth.shape[1]]
    inmask = dil[mr, mc]
    anchor = float(model_depth[model_depth.shape[0] // 2, model_depth.shape[1] // 2]) / 1000.0
    m_sel = inmask & (md > 0)
    z = md[m_sel].astype(np.float64) / 1000.0
    # (model_c - K02) / K00 * z : int - float -> float, / float -> float, * double -> double
    xm = ((mc[m_sel].astype(np.float32) - modelK[0, 2]) / modelK[0, 0]).astype(np.float64) * z
    ym = ((mr[m_sel].astype(np.float32) - modelK[1, 2]) / modelK[1, 1]).astype(np.float64) * z
    model_pts = np.stack([xm, ym, z], 1)
    sd = scene_depth[sr, sc]
    s_sel = inmask & (sd > 0)
    zs = sd.astype(np.float64) / 1000.0
    xs = ((sc.astype(np.float32) - sceneK[0, 2]) / sceneK[0, 0]).astype(np.float64) * zs
    ys = ((sr.astype(np.float32) - sceneK[1, 2]) / sceneK[1, 1]).astype(np.float64) * zs
    c_sel = s_sel & (np.abs(zs - anchor) < 0.4) & (md > 0)
    # sequential accumulation order of the reference (row-major); float64 sums
    center_scene = np.array([xs[c_sel].sum(), ys[c_sel].sum(), zs[c_sel].sum()])
    n_scene = int(c_sel.sum())
    center_model = model_pts.sum(0)
    with np.errstate(divide="ignore", invalid="ignore"):
        center_model = center_model / model_pts.shape[0]
        center_scene = center_scene / n_scene
    scene_pts = np.stack([xs[s_sel], ys[s_sel], zs[s_sel]], 1)
    return model_pts, center_model, center_scene, scene_pts


def voxel_down_sample(pts, voxel=VOXEL):
    """PointCloud::VoxelDownSample: mean of the points of every occupied voxel (sorted voxel order)."""
    if pts.shape[0] == 0:
        return pts
    origin = pts.min(0) - voxel * 0.5
    idx = np.floor((pts - origin) / voxel).astype(np.int64)
    order = np.lexsort((idx[:, 2], idx[:, 1], idx[:, 0]))
    idx, p = idx[order], pts[order]
    new = np.ones(len(idx), bool)
    new[1:] = np.any(idx[1:] != idx[:-1], 1)
    starts = np.nonzero(new)[0]
    sums = np.add.reduceat(p, starts, 0)
    counts = np.diff(np.append(starts, len(idx)))
    return sums / counts[:, None]


def estimate_normals(pts, knn=KNN):
    """EstimateNormals(KDTreeSearchParamKNN(30)): smallest eigenvector of the neighbourhood covariance."""
    n = pts.shape[0]
    normals = np.zeros((n, 3))
    if n == 0:
        return normals
    k = min(knn, n)
    _, nb = cKDTree(pts).query(pts, k=k)
    nb = nb.reshape(n, k)
    for i in range(n):
        if k < 3:
            normals[i] = (0, 0, 1)
            continue
        q = pts[nb[i]]
        mean = q.mean(0)
        cov = (q[:, :, None] * q[:, None, :]).mean(0) - np.outer(mean, mean)
        w, v = np.linalg.eigh(cov)
        nrm = v[:, 0]
        normals[i] = nrm if np.linalg.norm(nrm) > 0 else (0, 0, 1)
    return normals


def vec6_to_mat4(x):
    """TransformVector6dToMatrix4d: Rz(x2) * Ry(x1) * Rx(x0), translation x3..5."""
    cx, sx, cy, sy, cz, sz = np.cos(x[0]), np.sin(x[0]), np.cos(x[1]), np.sin(x[1]), np.cos(x[2]), np.sin(x[2])
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    T = np.eye(4)
    T[:3, :3] = Rz @ Ry @ Rx
    T[:3, 3] = x[3:]
    return T


def registration_icp_p2pl(src, tgt, tgt_normals, max_d, init, max_iter=MAX_ITER, rel_fitness=REL_FITNESS, rel_rmse=REL_RMSE):
    """RegistrationICP(..., TransformationEstimationPointToPlane()).  Returns (T 4x4, fitness, rmse, iterations)."""
    T = np.array(init, np.float64)
    if src.shape[0] == 0 or tgt.shape[0] == 0:
        return T, 0.0, 0.0, 0
    tree = cKDTree(tgt)
    p = src @ T[:3, :3].T + T[:3, 3]

    def correspond(p):
        d, j = tree.query(p, k=1)
        ok = d * d < max_d * max_d  # radius search is strict in FLANN
        n = int(ok.sum())
        fit = n / p.shape[0]
        rmse = float(np.sqrt((d[ok] ** 2).sum() / n)) if n else 0.0
        return np.nonzero(ok)[0], j[ok], fit, rmse

    si, tj, fit, rmse = correspond(p)
    it = 0
    for it in range(1, max_iter + 1):
        vs, vt, nt = p[si], tgt[tj], tgt_normals[tj]
        r = ((vs - vt) * nt).sum(1)
        J = np.concatenate([np.cross(vs, nt), nt], 1)
        JTJ = J.T @ J
        JTr = J.T @ r
        try:
            x = np.linalg.solve(JTJ, -JTr)
            upd = vec6_to_mat4(x)
            if not np.all(np.isfinite(upd)):
                upd = np.eye(4)
        except np.linalg.LinAlgError:
            upd = np.eye(4)
        T = upd @ T
        p = p @ upd[:3, :3].T + upd[:3, 3]
        fit0, rmse0 = fit, rmse
        si, tj, fit, rmse = correspond(p)
        if abs(fit0 - fit) < rel_fitness and abs(rmse0 - rmse) < rel_rmse:
            break
    return T, fit, rmse, it


def pose_refine(scene_depth, model_depth, sceneK, modelK, modelR, modelT, detectX, detectY, max_iter=MAX_ITER,
                use_scene_cloud=False):
    """poseRefine::process.  Returns dict(R 3x3 f64, t 3x1 f64 [mm], residual float, iterations) ;
    residual == -1 and R, t None on the early return (LL.cpp:52-55)."""
    scene_depth = np.asarray(scene_depth)
    model_depth = np.asarray(model_depth)
    init_base = np.zeros((4, 4), np.float32)
    init_base[:3, :3] = np.asarray(modelR, np.float32).reshape(3, 3)
    init_base[:3, 3] = np.asarray(modelT, np.float32).reshape(3)
    init_base[2, 3] = init_base[2, 3] / np.float32(1000.0)  # only the z component (LL.cpp:37)
    init_base[3, 3] = 1
    bbox = model_bbox(model_depth)
    if detectX + bbox[2] >= scene_depth.shape[1] or detectY + bbox[3] >= scene_depth.shape[0]:
        return dict(R=None, t=None, residual=-1.0, iterations=0)
    model_pts, c_model, c_scene, scene_pts = build_clouds(scene_depth, model_depth, sceneK, modelK, detectX, detectY, bbox)
    init_guess = np.eye(4)
    init_guess[:3, 3] = c_scene - c_model
    down = voxel_down_sample(model_pts)
    target = voxel_down_sample(scene_pts) if use_scene_cloud else down  # the reference: model cloud again (LL.cpp:109)
    normals = estimate_normals(target)
    T, fit, rmse, it = registration_icp_p2pl(down, target, normals, MAX_DIST, init_guess, max_iter)
    result = T @ init_base.astype(np.float64)
    return dict(R=result[:3, :3].copy(), t=(result[:3, 3] * 1000.0).reshape(3, 1), residual=float(np.float32(fit)), iterations=it,
                n_points=int(down.shape[0]), rmse=rmse)
