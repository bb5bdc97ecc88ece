"""Shared helpers for GPU-vs-oracle parity tests."""
import numpy as np

ROBOT_FIELDS = ["x", "y", "th", "v", "w", "wpx", "wpy", "pdist", "phead", "ppx", "ppy", "step", "episode",
                "flags", "cnt0", "cnt1"]
_F32_FIELDS = ("v", "w", "wpx", "wpy", "pdist", "phead", "ppx", "ppy")


def describe_blob_diff(cfg, a: np.ndarray, b: np.ndarray, max_lines: int = 12) -> str:
    """Human-readable first differences between two state blobs (a = GPU, b = oracle)."""
    E, N = cfg.n_envs, cfg.n_peds
    lines = []
    ra, rb = a[16:16 + E * 16].reshape(E, 16), b[16:16 + E * 16].reshape(E, 16)
    for e, f in zip(*np.nonzero(ra != rb)):
        va, vb = ra[e, f], rb[e, f]
        if ROBOT_FIELDS[f] in _F32_FIELDS:
            va, vb = va.view(np.float32), vb.view(np.float32)
        lines.append("robot env %d %s: gpu=%r oracle=%r" % (e, ROBOT_FIELDS[f], va, vb))
        if len(lines) >= max_lines:
            return "\n".join(lines)
    o = 16 + E * 16
    for name, off in (("ped_a", o), ("ped_b", o + E * N * 4)):
        pa, pb = a[off:off + E * N * 4].reshape(E, N, 4), b[off:off + E * N * 4].reshape(E, N, 4)
        for e, n, f in zip(*np.nonzero(pa != pb)):
            lines.append("%s env %d ped %d word %d: gpu=%#x oracle=%#x (f32 %r vs %r)" % (
                name, e, n, f, pa[e, n, f], pb[e, n, f], pa[e, n, f].view(np.float32), pb[e, n, f].view(np.float32)))
            if len(lines) >= max_lines:
                return "\n".join(lines)
    return "\n".join(lines)


def describe_obs_diff(cfg, a: np.ndarray, b: np.ndarray, max_lines: int = 12) -> str:
    NR = cfg.n_samples - 1
    names = ["ray%d" % j for j in range(NR)] + ["heading", "dist", "x", "y", "yaw", "avx", "avy"]
    names += ["blk%d.%s" % (s, c) for s in range(cfg.k_obstacles) for c in ("x", "y", "vx", "vy")]
    lines = []
    av, bv = a.view(np.uint32), b.view(np.uint32)
    for e, k in zip(*np.nonzero(av != bv)):
        lines.append("obs env %d %s: gpu=%r oracle=%r" % (e, names[k], a[e, k], b[e, k]))
        if len(lines) >= max_lines:
            break
    return "\n".join(lines)


def bits_equal(a: np.ndarray, b: np.ndarray) -> bool:
    return a.shape == b.shape and np.array_equal(np.ascontiguousarray(a).view(np.uint8),
                                                 np.ascontiguousarray(b).view(np.uint8))


def random_actions(rng: np.random.Generator, E: int) -> np.ndarray:
    """v ~ U(0, 0.22), w ~ U(-2, 2): the agents' action box (TD3DRV:67-68)."""
    a = np.empty((E, 2), dtype=np.float32)
    a[:, 0] = rng.uniform(0.0, 0.22, E)
    a[:, 1] = rng.uniform(-2.0, 2.0, E)
    return a
