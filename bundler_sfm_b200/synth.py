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
