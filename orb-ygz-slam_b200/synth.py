"""Deterministic synthetic inputs.  Integer arithmetic only where the result is an
image that fixtures depend on, so a seed gives the same bytes on every host.

  synth_frame(seed, w, h)          busy gray frame: band-limited noise + rectangles + a flat region
                                   + a low-contrast region (forces the 20 -> 7 FAST fallback)
  synth_stream(seed, n, w, h)      EuRoC-shape mono stream: crops of one 2x texture along a smooth walk
  synth_descriptors(...)           N x 32 random descriptors with planted matches
  synth_sia_pair(...)              ref/cur pair of a fronto-parallel textured plane with known pose
"""
import numpy as np


def _box3(img):
    p = np.pad(img.astype(np.uint32), 1, mode="edge")
    s = (p[:-2, :-2] + p[:-2, 1:-1] + p[:-2, 2:] + p[1:-1, :-2] + p[1:-1, 1:-1] + p[1:-1, 2:] +
         p[2:, :-2] + p[2:, 1:-1] + p[2:, 2:])
    return ((s + 4) // 9).astype(np.uint32)


def _stretch(img):
    lo, hi = int(img.min()), int(img.max())
    if hi == lo:
        return np.zeros(img.shape, np.uint8)
    return (((img.astype(np.int64) - lo) * 255 + (hi - lo) // 2) // (hi - lo)).astype(np.uint8)


def synth_texture(seed, w, h, passes=2, n_rect=150):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, size=(h, w), dtype=np.uint32)
    for _ in range(passes):
        img = _box3(img)
    img = _stretch(img)
    for _ in range(n_rect):
        rw, rh = int(rng.integers(8, 64)), int(rng.integers(8, 64))
        x0, y0 = int(rng.integers(0, max(1, w - rw))), int(rng.integers(0, max(1, h - rh)))
        g = int(rng.integers(0, 256))
        # keep some texture inside the rectangle so FAST scores vary
        patch = img[y0:y0 + rh, x0:x0 + rw].astype(np.int32)
        img[y0:y0 + rh, x0:x0 + rw] = np.clip(g + (patch - 128) // 4, 0, 255).astype(np.uint8)
    return img


def synth_frame(seed, w=752, h=480):
    img = synth_texture(seed, w, h)
    rng = np.random.default_rng(seed + 7919)
    # ~10 % flat (no corners at any threshold)
    fw, fh = int(w * 0.4), int(h * 0.25)
    fx, fy = int(rng.integers(0, w - fw)), int(rng.integers(0, h - fh))
    img[fy:fy + fh, fx:fx + fw] = 97
    # ~12 % low contrast (corners only at the fallback threshold)
    lw, lh = int(w * 0.35), int(h * 0.35)
    lx, ly = int(rng.integers(0, w - lw)), int(rng.integers(0, h - lh))
    reg = img[ly:ly + lh, lx:lx + lw].astype(np.int32)
    img[ly:ly + lh, lx:lx + lw] = np.clip(128 + (reg - 128) // 7, 0, 255).astype(np.uint8)
    return np.ascontiguousarray(img)


def synth_stream(seed, n, w=752, h=480):
    """n frames cropped from one (2w x 2h) texture along a smooth integer random walk."""
    tex = synth_frame(seed, 2 * w, 2 * h)
    rng = np.random.default_rng(seed + 104729)
    vx, vy = 3.0, 1.0
    x, y = w / 2.0, h / 2.0
    out = np.empty((n, h, w), np.uint8)
    for i in range(n):
        vx = 0.9 * vx + float(rng.normal()) * 1.5
        vy = 0.9 * vy + float(rng.normal()) * 1.0
        x = min(max(x + vx, 0), w - 1)
        y = min(max(y + vy, 0), h - 1)
        xi, yi = int(x), int(y)
        out[i] = tex[yi:yi + h, xi:xi + w]
    return out


def synth_descriptors(seed, nq, nt, planted=0.6, flip_p=0.08):
    """Random 256-bit descriptors; a `planted` fraction of queries are noisy copies of targets."""
    rng = np.random.default_rng(seed)
    t = rng.integers(0, 256, size=(nt, 32), dtype=np.uint8)
    q = rng.integers(0, 256, size=(nq, 32), dtype=np.uint8)
    truth = np.full(nq, -1, np.int32)
    n_pl = int(nq * planted)
    src = rng.permutation(nt)[:n_pl] if n_pl <= nt else rng.integers(0, nt, n_pl)
    for i, j in enumerate(src):
        bits = np.unpackbits(t[j])
        flips = rng.random(256) < flip_p
        q[i] = np.packbits(bits ^ flips.astype(np.uint8))
        truth[i] = j
    perm = rng.permutation(nq)
    return q[perm], t, truth[perm]


def quat_from_rotvec(rv):
    rv = np.asarray(rv, np.float64)
    th = np.linalg.norm(rv)
    if th < 1e-12:
        return np.array([0, 0, 0, 1], np.float32)
    ax = rv / th
    s = np.sin(th / 2)
    return np.array([ax[0] * s, ax[1] * s, ax[2] * s, np.cos(th / 2)], np.float32)


def quat_to_R(q):
    x, y, z, w = [float(v) for v in q]
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]], np.float64)


EUROC_K = (458.654, 457.296, 367.215, 248.375)   # Examples/Monocular/EuRoC.yaml:8-11


def synth_sia_pair(seed, n_pts=200, w=752, h=480, K=EUROC_K, depth=4.0, max_t=0.05, max_rot_deg=1.0):
    """ref/cur images of a fronto-parallel textured plane z=depth (in the ref camera frame), the
    ref camera at identity.  Returns dict(ref, cur, px[n,2], xyz[n,3] (world == ref cam), q_true, t_true)
    where (q_true, t_true) = T_cur_from_ref."""
    fx, fy, cx, cy = K
    rng = np.random.default_rng(seed)
    tex = synth_texture(seed + 31, w, h, passes=6, n_rect=40).astype(np.float64)
    rv = (rng.random(3) * 2 - 1) * np.deg2rad(max_rot_deg)
    t = (rng.random(3) * 2 - 1) * max_t
    q = quat_from_rotvec(rv)
    R = quat_to_R(q)                       # x_cur = R x_ref + t
    # render cur: ray through each cur pixel, expressed in ref frame, hit plane z = depth
    uu, vv = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
    d_cur = np.stack([(uu - cx) / fx, (vv - cy) / fy, np.ones_like(uu)], -1)
    Rt = R.T
    d_ref = d_cur @ Rt.T                   # R^T d
    o_ref = -Rt @ t                        # cur camera centre in ref frame
    lam = (depth - o_ref[2]) / d_ref[..., 2]
    X = o_ref[0] + lam * d_ref[..., 0]
    Y = o_ref[1] + lam * d_ref[..., 1]
    su = fx * X / depth + cx
    sv = fy * Y / depth + cy
    x0 = np.clip(np.floor(su).astype(np.int64), 0, w - 2)
    y0 = np.clip(np.floor(sv).astype(np.int64), 0, h - 2)
    ax = np.clip(su - x0, 0, 1)
    ay = np.clip(sv - y0, 0, 1)
    cur = ((1 - ax) * (1 - ay) * tex[y0, x0] + ax * (1 - ay) * tex[y0, x0 + 1] +
           (1 - ax) * ay * tex[y0 + 1, x0] + ax * ay * tex[y0 + 1, x0 + 1])
    cur = np.clip(np.rint(cur), 0, 255).astype(np.uint8)
    ref = tex.astype(np.uint8)
    px = np.stack([rng.uniform(40, w - 40, n_pts), rng.uniform(40, h - 40, n_pts)], -1).astype(np.float32)
    xyz = np.stack([(px[:, 0].astype(np.float64) - cx) / fx * depth,
                    (px[:, 1].astype(np.float64) - cy) / fy * depth,
                    np.full(n_pts, depth)], -1).astype(np.float32)
    return dict(ref=np.ascontiguousarray(ref), cur=np.ascontiguousarray(cur), px=px, xyz=xyz,
                q_true=q, t_true=t.astype(np.float32), K=K)


def render_plane(tex, K, q, t, depth, w, h):
    """Image of the textured world plane z = depth seen by a camera with pose x_cam = R(q) x_world + t.  The texture is
    pinned to the plane by su = fx X / depth + cx, sv = fy Y / depth + cy (so the identity camera sees `tex` itself)."""
    fx, fy, cx, cy = K
    R = quat_to_R(q)
    t = np.asarray(t, np.float64)
    uu, vv = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
    d_cam = np.stack([(uu - cx) / fx, (vv - cy) / fy, np.ones_like(uu)], -1)
    d_w = d_cam @ R                        # R^T d
    o_w = -R.T @ t                         # camera centre in the world
    lam = (depth - o_w[2]) / d_w[..., 2]
    X = o_w[0] + lam * d_w[..., 0]
    Y = o_w[1] + lam * d_w[..., 1]
    su = fx * X / depth + cx
    sv = fy * Y / depth + cy
    th, tw = tex.shape
    inside = (su >= 0) & (sv >= 0) & (su <= tw - 1) & (sv <= th - 1)
    x0 = np.clip(np.floor(su).astype(np.int64), 0, tw - 2)
    y0 = np.clip(np.floor(sv).astype(np.int64), 0, th - 2)
    ax = np.clip(su - x0, 0, 1)
    ay = np.clip(sv - y0, 0, 1)
    tex = tex.astype(np.float64)
    img = ((1 - ax) * (1 - ay) * tex[y0, x0] + ax * (1 - ay) * tex[y0, x0 + 1] +
           (1 - ax) * ay * tex[y0 + 1, x0] + ax * ay * tex[y0 + 1, x0 + 1])
    img = np.where(inside, img, 0.0)
    return np.ascontiguousarray(np.clip(np.rint(img), 0, 255).astype(np.uint8))


def synth_direct_scene(seed, n_kf=3, n_items=300, w=752, h=480, K=EUROC_K, depth=4.0, nlevels=8, scale=1.2, noise_px=1.5):
    """Keyframes + one current frame looking at the textured plane z = depth, and (map point, keyframe) items for
    ORBmatcher::FindDirectProjection.  The last keyframe stands twice as far from the plane (affine determinant ~4,
    so its items search one pyramid level up).  Returns dict(kf_imgs, kf_poses [(q,t)], cur_img, cur_pose, items
    (structured: ref, px_ref, octave, pw, px_cur), px_true [n,2])."""
    fx, fy, cx, cy = K
    rng = np.random.default_rng(seed)
    tex = synth_texture(seed + 77, w, h, passes=5, n_rect=60)
    def pose(max_t, max_rot_deg, tz=0.0):
        rv = (rng.random(3) * 2 - 1) * np.deg2rad(max_rot_deg)
        t = (rng.random(3) * 2 - 1) * max_t
        t[2] += tz
        return quat_from_rotvec(rv), t.astype(np.float32)
    kf_poses = [pose(0.08, 1.5) for _ in range(n_kf - 1)] + [pose(0.05, 1.0, tz=depth)]   # x_cam.z = z_w + t_z: twice as far
    cur_pose = pose(0.06, 1.0)
    kf_imgs = [render_plane(tex, K, q, t, depth, w, h) for q, t in kf_poses]
    cur_img = render_plane(tex, K, cur_pose[0], cur_pose[1], depth, w, h)
    dt = np.dtype([("ref", np.int32), ("px_ref", np.float32, 2), ("octave", np.int32), ("pw", np.float32, 3),
                   ("px_cur", np.float32, 2)])
    items = np.zeros(n_items, dt)
    px_true = np.zeros((n_items, 2), np.float32)
    Rc = quat_to_R(cur_pose[0]); tc = np.asarray(cur_pose[1], np.float64)
    for i in range(n_items):
        k = int(rng.integers(0, n_kf))
        q, t = kf_poses[k]
        R = quat_to_R(q); t = np.asarray(t, np.float64)
        margin = 8 if i % 11 == 0 else 40            # a few items hug the border: warp samples fall outside -> zeros
        px = np.array([rng.uniform(margin, w - margin), rng.uniform(margin, h - margin)])
        d_w = R.T @ np.array([(px[0] - cx) / fx, (px[1] - cy) / fy, 1.0])
        o_w = -R.T @ t
        lam = (depth - o_w[2]) / d_w[2]
        pw = o_w + lam * d_w
        pc = Rc @ pw + tc
        uv = np.array([fx * pc[0] / pc[2] + cx, fy * pc[1] / pc[2] + cy])
        items["ref"][i] = k
        items["px_ref"][i] = px
        items["octave"][i] = int(rng.integers(0, min(nlevels, 4)))
        items["pw"][i] = pw
        items["px_cur"][i] = uv + (rng.random(2) * 2 - 1) * noise_px
        px_true[i] = uv
    return dict(kf_imgs=kf_imgs, kf_poses=kf_poses, cur_img=cur_img, cur_pose=cur_pose, items=items, px_true=px_true,
                K=K, nlevels=nlevels, scale=scale)


def synth_stereo_pair(seed, w=752, h=480, disparities=(6, 14, 25), noise=2):
    """Rectified pair: the right image is the left one shifted by a per-band disparity (uR = uL - d), plus small
    integer noise.  Returns (left, right, band_of_row[h], disparities)."""
    left = synth_frame(seed, w + max(disparities) + 4, h)
    rng = np.random.default_rng(seed + 15485863)
    right = np.empty((h, w), np.uint8)
    band = np.minimum(np.arange(h) * len(disparities) // h, len(disparities) - 1)
    for y in range(h):
        d = disparities[band[y]]
        right[y] = left[y, d:d + w]
    if noise:
        right = np.clip(right.astype(np.int32) + rng.integers(-noise, noise + 1, right.shape), 0, 255).astype(np.uint8)
    return np.ascontiguousarray(left[:, :w]), np.ascontiguousarray(right), band, tuple(disparities)


def synth_vocabulary(seed, k=10, L=3, stop_frac=0.05, early_leaf_frac=0.03):
    """A k-ary, depth-L ORB vocabulary in the binary layout of ORBvoc.bin (TemplatedVocabulary::saveToBinaryFile,
    Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1527-1551): header {u32 nb_nodes, u32 size_node = 41, i32 k, i32 L,
    i32 scoring = 0 (L1), i32 weighting = 0 (TF-IDF)} + one 41-byte record per node {i32 parent, u8 desc[32], f32 weight,
    u8 is_leaf}, level by level.  Children are bit-flipped copies of their parent, a few leaves sit above the last level,
    a few words are stopped (weight 0).  Returns bytes."""
    rng = np.random.default_rng(seed)
    rec = np.dtype([("parent", "<i4"), ("desc", "u1", 32), ("weight", "<f4"), ("leaf", "u1")])
    assert rec.itemsize == 41
    levels = []
    parent_ids = np.zeros(1, np.int64)                       # node ids of the current frontier (root = 0)
    parent_desc = rng.integers(0, 256, (1, 32), dtype=np.uint8)
    next_id = 1
    for level in range(1, L + 1):
        m = len(parent_ids) * k
        par = np.repeat(parent_ids, k)
        base = np.repeat(parent_desc, k, axis=0)
        flips = np.packbits(rng.random((m, 256)) < (0.22 / level), axis=1)
        d = base ^ flips
        leaf = np.full(m, level == L)
        if level >= 2 and level < L:
            leaf |= rng.random(m) < early_leaf_frac
        w = np.where(leaf, np.where(rng.random(m) < stop_frac, 0.0, rng.uniform(0.5, 9.0, m)), 0.0).astype(np.float32)
        r = np.zeros(m, rec)
        r["parent"], r["desc"], r["weight"], r["leaf"] = par, d, w, leaf
        levels.append(r)
        ids = np.arange(next_id, next_id + m)
        next_id += m
        parent_ids, parent_desc = ids[~leaf], d[~leaf]
        if len(parent_ids) == 0:
            break
    recs = np.concatenate(levels)
    return np.array([len(recs) + 1, 41], np.uint32).tobytes() + np.array([k, L, 0, 0], np.int32).tobytes() + recs.tobytes()
