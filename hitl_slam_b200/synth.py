"""Deterministic synthetic 2-D lidar trajectories in the reference's `.stfs.covars` format.

The example bags of ut-amrl/hitl-slam are not available offline, so the BASELINE.json configs are
realised here: a closed rectilinear wall map (outer box + a grid of solid
blocks, i.e. a lattice of corridors), a tour through the corridors with 0.25 m pose spacing
and tangent heading, a 270-degree lidar with range noise, wall normals facing the sensor, and
random-walk odometry drift so that revisits visibly misalign.  The scans are written with the
reference's writer format (4 decimals) and re-read through the loader mirror, so the inputs
carry the text quantisation and the loader's normal-translation quirk.

Randomness: counter-based SplitMix64 -> Box-Muller (no dependence on numpy's generators).
"""
import os

import numpy as np

from .capi import HostLib

CONFIGS = {
    # name: (blocks_x, blocks_y, block_w, block_h, corridor, n_poses, beams)
    "c1": dict(bx=2, by=1, bw=9.0, bh=4.0, cor=2.0, n_poses=500, beams=360, seed=0xC0FFEE + 1),     # figure-8
    "c2": dict(bx=6, by=1, bw=8.0, bh=6.0, cor=2.0, n_poses=5000, beams=720, seed=0xC0FFEE + 2),    # corridor + rooms
    "c3": dict(bx=2, by=2, bw=20.0, bh=14.0, cor=3.0, n_poses=20000, beams=1080, seed=0xC0FFEE + 3),  # campus loops
    "c4": dict(bx=4, by=2, bw=10.0, bh=6.0, cor=2.0, n_poses=10000, beams=720, seed=0xC0FFEE + 4),
    "tiny": dict(bx=1, by=1, bw=4.0, bh=3.0, cor=1.5, n_poses=40, beams=96, seed=0xC0FFEE + 9),
    "small": dict(bx=2, by=1, bw=5.0, bh=3.0, cor=1.5, n_poses=160, beams=180, seed=0xC0FFEE + 10),
}


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def _uniform(seed, stream, n):
    with np.errstate(over="ignore"):
        idx = np.arange(n, dtype=np.uint64) + np.uint64(seed) * np.uint64(0x100000001B3) + np.uint64(stream) * np.uint64(0x9E3779B1)
        bits = _splitmix64(idx)
    return ((bits >> np.uint64(11)).astype(np.float64) + 0.5) / float(1 << 53)


def _normal(seed, stream, n):
    u1, u2 = _uniform(seed, 2 * stream, n), _uniform(seed, 2 * stream + 1, n)
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


def make_world(bx, by, bw, bh, cor):
    """Axis-aligned wall segments (x0, y0, x1, y1) of the outer box and the solid blocks, plus the
    corridor-centre lattice coordinates."""
    W = bx * bw + (bx + 1) * cor
    H = by * bh + (by + 1) * cor
    walls = [(0, 0, W, 0), (W, 0, W, H), (W, H, 0, H), (0, H, 0, 0)]
    for i in range(bx):
        for j in range(by):
            x0, y0 = cor + i * (bw + cor), cor + j * (bh + cor)
            x1, y1 = x0 + bw, y0 + bh
            walls += [(x0, y0, x1, y0), (x1, y0, x1, y1), (x1, y1, x0, y1), (x0, y1, x0, y0)]
    xs = [cor / 2 + i * (bw + cor) for i in range(bx + 1)]
    ys = [cor / 2 + j * (bh + cor) for j in range(by + 1)]
    return np.array(walls, np.float64), xs, ys


def make_tour(bx, by, xs, ys):
    """Closed tour over the corridor lattice: circle every block once, row by row."""
    pts = []
    for j in range(by):
        cols = range(bx)
        for i in cols:
            ll, lr, ur, ul = (xs[i], ys[j]), (xs[i + 1], ys[j]), (xs[i + 1], ys[j + 1]), (xs[i], ys[j + 1])
            pts += [ll, lr, ur, ul, ll, lr]
        # climb on the right edge, return along the top of this row to the left edge
        pts += [(xs[bx], ys[j + 1]), (xs[0], ys[j + 1])]
    # back down the left edge to the start
    pts += [(xs[0], ys[0])]
    out = [pts[0]]
    for p in pts[1:]:
        if p != out[-1]:
            out.append(p)
    return np.array(out, np.float64)


def sample_path(tour, n_poses, spacing=0.25):
    seg = np.diff(tour, axis=0)
    seglen = np.hypot(seg[:, 0], seg[:, 1])
    cum = np.concatenate([[0.0], np.cumsum(seglen)])
    total = cum[-1]
    s = (np.arange(n_poses) * spacing) % total
    k = np.clip(np.searchsorted(cum, s, side="right") - 1, 0, len(seg) - 1)
    t = (s - cum[k]) / seglen[k]
    xy = tour[k] + seg[k] * t[:, None]
    th = np.arctan2(seg[k, 1], seg[k, 0])
    return xy, th


def raycast(xy, th, beams, walls, fov=np.deg2rad(270.0), chunk=256):
    """Nearest wall hit per beam. Returns range [N, beams], wall id [N, beams]."""
    n = len(xy)
    ang_rel = np.linspace(-fov / 2, fov / 2, beams)
    rng = np.full((n, beams), np.inf)
    wid = np.zeros((n, beams), np.int32)
    vert = np.where(walls[:, 0] == walls[:, 2])[0]
    hori = np.where(walls[:, 1] == walls[:, 3])[0]
    for a in range(0, n, chunk):
        b = min(n, a + chunk)
        ang = th[a:b, None] + ang_rel[None, :]
        dx, dy = np.cos(ang), np.sin(ang)
        ox, oy = xy[a:b, 0, None], xy[a:b, 1, None]
        best = np.full(ang.shape, np.inf)
        bid = np.zeros(ang.shape, np.int32)
        with np.errstate(divide="ignore", invalid="ignore"):
            for w in vert:
                x0, y0, _, y1 = walls[w]
                t = (x0 - ox) / dx
                y = oy + t * dy
                ok = (t > 1e-9) & (y >= min(y0, y1)) & (y <= max(y0, y1)) & (t < best)
                best = np.where(ok, t, best)
                bid = np.where(ok, w, bid)
            for w in hori:
                x0, y0, x1, _ = walls[w]
                t = (y0 - oy) / dy
                x = ox + t * dx
                ok = (t > 1e-9) & (x >= min(x0, x1)) & (x <= max(x0, x1)) & (t < best)
                best = np.where(ok, t, best)
                bid = np.where(ok, w, bid)
        rng[a:b] = best
        wid[a:b] = bid
    return rng, wid, ang_rel


def generate(name="c1", n_poses=None, beams=None, normals="compensated", max_range=30.0, range_sigma=0.01, drift_xy=0.005,
             drift_th=0.002, out_dir=None, keep_file=False, **override):
    """Returns dict(poses [N,3] f32, cov [N,9] f32, offsets [N+1] u32, pts [M,2] f32, nrm [M,2] f32, path)
    after the text round trip through the `.stfs.covars` format."""
    cfg = dict(CONFIGS[name])
    cfg.update(override)
    if n_poses is not None:
        cfg["n_poses"] = n_poses
    if beams is not None:
        cfg["beams"] = beams
    N, P, seed = cfg["n_poses"], cfg["beams"], cfg["seed"]
    walls, xs, ys = make_world(cfg["bx"], cfg["by"], cfg["bw"], cfg["bh"], cfg["cor"])
    tour = make_tour(cfg["bx"], cfg["by"], xs, ys)
    xy_true, th_true = sample_path(tour, N)
    # small lateral wobble so revisits are not perfectly identical
    xy_true = xy_true + 0.05 * np.stack([_normal(seed, 1, N), _normal(seed, 2, N)], 1)
    th_true = th_true + 0.01 * _normal(seed, 3, N)
    rng, wid, ang_rel = raycast(xy_true, th_true, P, walls)
    noise = range_sigma * _normal(seed, 4, N * P).reshape(N, P)
    valid = np.isfinite(rng) & (rng <= max_range)
    r = rng + noise
    # robot-frame points and normals
    px, py = r * np.cos(ang_rel)[None, :], r * np.sin(ang_rel)[None, :]
    wdx, wdy = walls[wid, 2] - walls[wid, 0], walls[wid, 3] - walls[wid, 1]
    wl = np.hypot(wdx, wdy)
    nxw, nyw = -wdy / wl, wdx / wl
    ang = th_true[:, None] + ang_rel[None, :]
    flip = (nxw * np.cos(ang) + nyw * np.sin(ang)) > 0   # make the normal face the sensor
    nxw, nyw = np.where(flip, -nxw, nxw), np.where(flip, -nyw, nyw)
    c, s = np.cos(-th_true)[:, None], np.sin(-th_true)[:, None]
    nxr, nyr = c * nxw - s * nyw, s * nxw + c * nyw
    # estimated (drifted) poses: random-walk error accumulated along the trajectory
    ex = np.cumsum(drift_xy * _normal(seed, 5, N))
    ey = np.cumsum(drift_xy * _normal(seed, 6, N))
    eth = np.cumsum(drift_th * _normal(seed, 7, N))
    xe, ye, the = xy_true[:, 0] + ex, xy_true[:, 1] + ey, th_true + eth
    the = np.arctan2(np.sin(the), np.cos(the))
    ce, se = np.cos(the)[:, None], np.sin(the)[:, None]
    owx, owy = ce * px - se * py + xe[:, None], se * px + ce * py + ye[:, None]
    nwx, nwy = ce * nxr - se * nyr, se * nxr + ce * nyr
    if normals == "compensated":   # loader does R(-th)(n - t): store n + t so the loaded normal is unit
        nwx, nwy = nwx + xe[:, None], nwy + ye[:, None]
    elif normals != "faithful":
        raise ValueError(normals)
    counts = valid.sum(1)
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint32)
    sel = valid.reshape(-1)
    obs = np.stack([owx.reshape(-1)[sel], owy.reshape(-1)[sel]], 1).astype(np.float32)
    nrm = np.stack([nwx.reshape(-1)[sel], nwy.reshape(-1)[sel]], 1).astype(np.float32)
    poses = np.stack([xe, ye, the], 1).astype(np.float32)
    i = np.arange(N, dtype=np.float64)
    cov = np.zeros((N, 9), np.float32)
    cov[:, 0] = cov[:, 4] = 1e-4 * (1 + i / 100)
    cov[:, 8] = 1e-5 * (1 + i / 100)
    host = HostLib()
    out_dir = out_dir or os.environ.get("HITL_SYNTH_DIR", "/tmp/hitl_synth")
    os.makedirs(out_dir, exist_ok=True)
    # scratch file of the text round trip: unique per process (ranks of one box generate the same map side by side)
    path = os.path.join(out_dir, "%s_%d_%d_%s%s.stfs.covars" % (name, N, P, normals, "" if keep_file else "_%d" % os.getpid()))
    host.save_stfs_covars(path, poses, cov, off, obs, nrm, map_name="synthetic_" + name, timestamp=0.0)
    g = host.load_pose_graph(path)
    g["path"] = path
    g["config"] = dict(cfg, name=name, normals=normals)
    if not keep_file:
        try:
            os.remove(path)
        except OSError:
            pass
    return g


def make_strokes(g, kind="colinear"):
    """Two strokes (feature A on a late visit, feature B on an early visit of the same wall) picked from
    the generated map: returns 4 points (a0, a1, b0, b1) in world coordinates as float32 [4,2]."""
    cfg = g["config"]
    cor = cfg["cor"]
    # bottom wall of block (0,0): y = cor, x in [cor, cor + bw]; strokes along it, slightly offset
    x0 = cor + 0.2 * cfg["bw"]
    x1 = cor + 0.6 * cfg["bw"]
    a = [(x0, cor + 0.01), (x1, cor + 0.012)]
    b = [(x0 + 0.05, cor - 0.005), (x1 - 0.1, cor - 0.004)]
    return np.array(a + b, np.float32)


def world_points(g):
    """World-frame points (float64, for picking strokes only) and the pose index of every point."""
    poses = g["poses"].astype(np.float64)
    off = g["offsets"].astype(np.int64)
    pid = np.repeat(np.arange(len(poses)), np.diff(off))
    c, s = np.cos(poses[pid, 2]), np.sin(poses[pid, 2])
    p = g["pts"].astype(np.float64)
    return np.stack([c * p[:, 0] - s * p[:, 1] + poses[pid, 0], s * p[:, 0] + c * p[:, 1] + poses[pid, 1]], 1), pid


def _observers(w, pid, a, b, thr=0.03, min_obs=5):
    """Poses with more than min_obs points inside the pill around segment a-b (EstablishObservationSets, approximately)."""
    d = b - a
    t = np.clip(((w - a) @ d) / (d @ d), 0.0, 1.0)
    dist = np.linalg.norm(w - (a + t[:, None] * d), axis=1)
    ids, cnt = np.unique(pid[dist < thr], return_counts=True)
    return ids[cnt > min_obs]


def _wall_windows(cfg, window):
    """Every (wall y, x0, x1) window a stroke pair may be drawn on: the horizontal walls of every block."""
    cor, bw, bh = cfg["cor"], cfg["bw"], cfg["bh"]
    out = []
    for j in range(cfg["by"]):
        for i in range(cfg["bx"]):
            for ywall in (cor + j * (bh + cor), cor + j * (bh + cor) + bh):
                wx0, wx1 = cor + i * (bw + cor) + 0.1 * bw, cor + i * (bw + cor) + 0.9 * bw
                for x0 in np.arange(wx0, wx1 - window + 1e-9, 0.5):
                    out.append((ywall, float(x0), float(x0 + window)))
    return out


def pick_strokes(g, min_sep=0.07, max_sep=0.6, window=2.0, start=0, return_next=False):
    """Two strokes a human would draw on the displayed map to close a loop: the same physical wall as
    it appears on a LATE visit (feature A, first stroke) and on an EARLY visit (feature B), picked where
    odometry drift separates the two appearances by more than the EM pill and the poses observing the
    two strokes are cleanly ordered in time.  Returns float32 [4, 2] (a0, a1, b0, b1) or raises if the
    map has no usable revisit.  `start` rotates the order in which the walls are tried (a replay of many
    corrections draws on a different place each time); with return_next the index to continue from is returned too."""
    cfg = g["config"]
    w, pid = world_points(g)
    wins = _wall_windows(cfg, window)
    w_all, pid_all, band_of = w, pid, None
    for q in range(len(wins)):
        ywall, x0, x1 = wins[(start + q) % len(wins)]
        if band_of != ywall:                                   # points near this wall, once per wall (the map has millions of points)
            near = np.flatnonzero(np.abs(w_all[:, 1] - ywall) < 0.35)
            w, pid, band_of = w_all[near], pid_all[near], ywall
        sel = (np.abs(w[:, 1] - ywall) < 0.3) & (w[:, 0] > x0) & (w[:, 0] < x1)
        ids = np.unique(pid[sel])
        if len(ids) < 12:
            continue
        passes = [p for p in np.split(ids, np.where(np.diff(ids) > 25)[0] + 1) if len(p) >= 6]
        inwin = np.flatnonzero((w[:, 0] > x0 - 0.05) & (w[:, 0] < x1 + 0.05))     # a stroke's pill lies inside the window
        ww, wp = w[inwin], pid[inwin]
        tried = 0
        fits = []
        for p in passes:
            m = sel & np.isin(pid, p)
            fits.append(np.polyfit(w[m, 0], w[m, 1], 1) if m.sum() >= 40 else None)
        for late in range(len(passes) - 1, 0, -1):
            if tried > 24:
                break
            for early in range(late):
                if fits[late] is None or fits[early] is None:
                    continue
                (k1, c1), (k0, c0) = fits[late], fits[early]
                xm = 0.5 * (x0 + x1)
                sep = abs((k1 * xm + c1) - (k0 * xm + c0))
                if not (min_sep <= sep <= max_sep):
                    continue
                A = np.array([[x0, k1 * x0 + c1], [x1, k1 * x1 + c1]])
                B = np.array([[x0, k0 * x0 + c0], [x1, k0 * x1 + c0]])
                tried += 1
                if tried > 24:                                   # enough attempts on this window: move on
                    break
                fa, fb = _observers(ww, wp, A[0], A[1]), _observers(ww, wp, B[0], B[1])
                both = np.intersect1d(fa, fb)
                fa, fb = np.setdiff1d(fa, both), np.setdiff1d(fb, both)
                if len(fa) >= 3 and len(fb) >= 3 and fa.min() > fb.max() + 5:
                    out = np.concatenate([A, B]).astype(np.float32)
                    return (out, (start + q + 1) % len(wins)) if return_next else out
    raise RuntimeError("no revisited wall with enough drift in this map")
