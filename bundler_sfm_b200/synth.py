"""Synthetic workloads of SURVEY.md section 8(d): SIFT-like uint8 descriptors (MATCH, config 4) and
ring-camera bundle-adjustment scenes (BA, configs 2/3).  Deterministic given the seed."""
import numpy as np


def sift_like_descriptors(num_images, keys_per_image, seed=7, copy_frac=0.3, noise=0.08):
    """List of [K,128] uint8 arrays.  gamma(0.6) magnitudes -> L2 normalise -> clip 0.2 -> renormalise
    -> x512 -> floor -> clamp 255.  `copy_frac` of every image's keys are noisy copies of keys of the
    previous image so that true matches exist."""
    rng = np.random.default_rng(seed)
    if np.isscalar(keys_per_image):
        keys_per_image = [int(keys_per_image)] * num_images

    def quantise(f):
        f = f / np.maximum(np.linalg.norm(f, axis=1, keepdims=True), 1e-12)
        f = np.minimum(f, 0.2)
        f = f / np.maximum(np.linalg.norm(f, axis=1, keepdims=True), 1e-12)
        return np.clip(np.floor(f * 512.0), 0, 255).astype(np.uint8)

    out, prev_f = [], None
    for i in range(num_images):
        K = keys_per_image[i]
        f = rng.gamma(0.6, 1.0, size=(K, 128)).astype(np.float32)
        f = f / np.maximum(np.linalg.norm(f, axis=1, keepdims=True), 1e-12)
        if prev_f is not None and K > 0 and prev_f.shape[0] > 0:
            nc = min(int(copy_frac * K), prev_f.shape[0])
            src = rng.choice(prev_f.shape[0], size=nc, replace=False)
            dst = rng.choice(K, size=nc, replace=False)
            f[dst] = np.abs(prev_f[src] + noise * rng.standard_normal((nc, 128)).astype(np.float32) * np.linalg.norm(prev_f[src], axis=1, keepdims=True) / np.sqrt(128.0))
        prev_f = f
        out.append(quantise(f))
    return out


def random_descriptors(n, seed, lo=0, hi=256):
    """uniform uint8 descriptors (adversarial/edge-case tests)"""
    rng = np.random.default_rng(seed)
    return rng.integers(lo, hi, size=(n, 128), dtype=np.int64).astype(np.uint8)


# ------------------------------------------------------------------------------------------------
# BA scenes: cameras on a ring of radius 6 looking at the origin (Bundler convention:
# the camera looks down -z, P = R (X - c)), f = 800, k = 0; points uniform in [-1,1]^3; each point is
# seen by L cameras; pixel noise N(0, 0.5^2); perturbed initial estimate.
# ------------------------------------------------------------------------------------------------
def ba_scene(num_cameras=50, num_points=20000, views_per_point=5, seed=1234, pixel_noise=0.5,
             pt_noise=0.02, cam_noise=0.01, f_noise=0.01, focal=800.0):
    """Returns a dict with the arguments of run_sfm (numpy arrays):
    vmask [n,m] int8, projections [nvis,2], R [m,9], c [m,3] (camera centres), f [m], k [m,2], pts [n,3]
    (the perturbed initial estimate) and the ground truth (gt_*)."""
    rng = np.random.default_rng(seed)
    m, n, L = num_cameras, num_points, views_per_point
    ang = 2.0 * np.pi * np.arange(m) / m
    c = np.stack([6.0 * np.cos(ang), 0.3 * np.sin(3 * ang), 6.0 * np.sin(ang)], 1)
    R = np.zeros((m, 3, 3))
    for j in range(m):
        z = c[j] / np.linalg.norm(c[j])          # camera looks down -z => z axis points away from the scene
        up = np.array([0.0, 1.0, 0.0])
        xax = np.cross(up, z); xax /= np.linalg.norm(xax)
        yax = np.cross(z, xax)
        R[j] = np.stack([xax, yax, z], 0)
    pts = rng.uniform(-1.0, 1.0, size=(n, 3))
    vmask = np.zeros((n, m), np.int8)
    step = max(m // L, 1)
    starts = (np.arange(n) * 7) % m
    for k in range(L):
        vmask[np.arange(n), (starts + k * step) % m] = 1
    pi, cj = np.nonzero(vmask)                    # row-major: point-major, ascending camera
    Pc = np.einsum("oij,oj->oi", R[cj], pts[pi] - c[cj])
    proj = -Pc[:, :2] * focal / Pc[:, 2:3]
    proj = proj + pixel_noise * rng.standard_normal(proj.shape)
    scene = {
        "vmask": vmask, "projections": np.ascontiguousarray(proj),
        "gt_pts": pts, "gt_c": c, "gt_R": R.reshape(m, 9).copy(), "gt_f": np.full(m, focal),
        "R": R.reshape(m, 9).copy(),
        "c": c + cam_noise * rng.standard_normal(c.shape),
        "f": focal * (1.0 + f_noise * rng.standard_normal(m)),
        "k": np.zeros((m, 2)),
        "pts": pts + pt_noise * rng.standard_normal(pts.shape),
    }
    return scene


def write_key_file(path, desc, seed=0):
    """Lowe text .key file (src/keys2a.cpp:183-253): "<num> 128", then per key "<y> <x> <scale> <ori>" and the
    128 descriptor values in 7 lines of 20,20,20,20,20,20,8 integers."""
    rng = np.random.default_rng(seed)
    n = desc.shape[0]
    with open(path, "w") as f:
        f.write(f"{n} 128\n")
        loc = rng.uniform(0, 1000, size=(n, 2)); sc = rng.uniform(1, 8, size=n); ori = rng.uniform(-3.14, 3.14, size=n)
        for i in range(n):
            f.write(f"{loc[i, 0]:.2f} {loc[i, 1]:.2f} {sc[i]:.2f} {ori[i]:.3f}\n")
            d = desc[i]
            for a in range(0, 128, 20):
                f.write(" " + " ".join(str(int(v)) for v in d[a:a + 20]) + "\n")
