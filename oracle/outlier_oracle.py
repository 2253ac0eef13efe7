"""oracle/outlier_oracle.py -- TEST INFRASTRUCTURE ONLY (never imported by the product).

CPU restatement of the reprojection statistics / outlier pass of BundlerApp::RunSFM_SBA (src/Bundle.cpp:659-856) on
run_sfm-shaped inputs.  The loop itself lives in application code, so it is restated here; its two building blocks are
the UNMODIFIED reference functions when oracle/_ref/libref_sba.so is present:
    sfm_project_rd      lib/sfm-driver/sfm.c:302-380        (projection with radial distortion, explicit centres)
    kth_element_copy    lib/imagelib/qsort.c:152-204         (k-th smallest; prints an error and returns 0.0 if k >= n)
and numpy restatements of those two otherwise ("port").  Pure-Python loops: small cases only."""
import ctypes

import numpy as np


def iround(x):                      # lib/imagelib/util.c:75-81
    return int(x - 0.5) if x < 0.0 else int(x + 0.5)


def _project_port(R, t, f, k, b, undistort):
    """sfm.c:326-377 with explicit_camera_centers = 1 and known_intrinsics = 0, same operation order"""
    b2 = np.array([b[0] - t[0], b[1] - t[1], b[2] - t[2]])
    bc = [R[0] * b2[0] + R[1] * b2[1] + R[2] * b2[2], R[3] * b2[0] + R[4] * b2[1] + R[5] * b2[2], R[6] * b2[0] + R[7] * b2[1] + R[8] * b2[2]]
    p0 = -bc[0] * f / bc[2]
    p1 = -bc[1] * f / bc[2]
    if undistort:
        rsq = (p0 * p0 + p1 * p1) / (f * f)
        factor = 1.0 + k[0] * rsq + k[1] * rsq * rsq
        p0 *= factor
        p1 *= factor
    return p0, p1


def _kth_port(vals, k):
    n = len(vals)
    if k >= n:
        return 0.0                  # qsort.c:192-194
    return float(np.sort(np.asarray(vals, dtype=np.float64))[k])


def reprojection_outliers(scene, estimate_distortion=1, min_thresh=8.0, max_thresh=16.0, pt_protected=None, use_reference=True):
    """scene: dict with vmask (n x m), projections (nvis x 2, point-major), R (m x 9), c (m x 3), f (m), k (m x 2), pts (n x 3).
    Returns dict(dist[nvis], stats[m x 5] = (n, mean, median, med80, thresh), outliers[], errors[], global_mean)."""
    from bundler_sfm_b200 import bundle
    lib = None
    if use_reference:
        from oracle import loader
        lib = loader.ref_sba()
    vmask = np.asarray(scene["vmask"]) != 0
    n, m = vmask.shape
    proj = np.asarray(scene["projections"], dtype=np.float64)
    cams = bundle.make_cameras(scene["R"], scene["c"], scene["f"], scene["k"])
    pts = np.ascontiguousarray(scene["pts"], dtype=np.float64)
    obs_pt, obs_cam = np.nonzero(vmask)              # row-major = the order of `projections`
    nvis = len(obs_pt)
    dist = np.zeros(nvis)
    if lib is not None:
        fn = lib.sfm_project_rd
        fn.restype = ctypes.c_int
        fn.argtypes = [ctypes.c_void_p] * 7 + [ctypes.c_int, ctypes.c_int]
        kth = lib.kth_element_copy
        kth.restype = ctypes.c_double
        kth.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    cam_size = ctypes.sizeof(bundle.CameraParams)
    base = ctypes.addressof(cams)
    for o in range(nvis):
        i, j = int(obs_pt[o]), int(obs_cam[o])
        if lib is not None:
            cam = bundle.CameraParams.from_address(base + j * cam_size)
            K = (ctypes.c_double * 9)(cam.f, 0, 0, 0, cam.f, 0, 0, 0, 1)                     # Bundle.cpp:670-672
            dt = (ctypes.c_double * 3)(cam.t[0], cam.t[1], cam.t[2])                        # :674-676
            b = (ctypes.c_double * 3)(*pts[i])
            pr = (ctypes.c_double * 2)()
            fn(base + j * cam_size, K, ctypes.addressof(cam) + bundle.CameraParams.k.offset, ctypes.addressof(cam), dt, b, pr,
               int(estimate_distortion), 1)                                                   # :741-744
            p0, p1 = pr[0], pr[1]
        else:
            p0, p1 = _project_port(scene["R"][j].reshape(-1), scene["c"][j], float(scene["f"][j]), scene["k"][j], pts[i], estimate_distortion)
        dx, dy = p0 - proj[o, 0], p1 - proj[o, 1]
        dist[o] = np.sqrt(dx * dx + dy * dy)                                                  # :750-753
    stats = np.zeros((m, 5))
    outliers, errors, first_cam = [], [], []
    tot, cnt = 0.0, 0
    for j in range(m):
        sel = np.nonzero(obs_cam == j)[0]             # ascending point index
        d = np.ascontiguousarray(dist[sel])
        nj = len(d)
        k80, k50 = iround(0.8 * nj), iround(0.5 * nj)
        if lib is not None and nj > 0:
            med80 = kth(nj, k80, d.ctypes.data) if k80 < nj else 0.0
            med50 = kth(nj, k50, d.ctypes.data) if k50 < nj else 0.0
        else:
            med80, med50 = _kth_port(d, k80), _kth_port(d, k50)
        thresh = 1.2 * 2.0 * med80                                                            # :767-768
        thresh = min_thresh if thresh < min_thresh else (max_thresh if thresh > max_thresh else thresh)   # CLAMP :769-771
        s = 0.0
        for v in d:                                                                           # :776-779
            s += v
        stats[j] = (nj, s / nj if nj else np.nan, med50, med80, thresh)
        tot += s; cnt += nj
        for q, o in enumerate(sel):                                                           # :793-823
            i = int(obs_pt[o])
            if pt_protected is not None and pt_protected[i]:
                continue
            if d[q] > thresh and i not in outliers:     # first camera that flags the point keeps its error (:809-821)
                outliers.append(i)
                errors.append(float(d[q]))
                first_cam.append(j)
    # the reference appends in (camera, key index) order; the key order inside a camera is application data that
    # run_sfm's arguments do not carry, so the list is reported by (first flagging camera, point index): same set, same errors
    order = sorted(range(len(outliers)), key=lambda q: (first_cam[q], outliers[q]))
    return {"dist": dist, "stats": stats, "outliers": np.array([outliers[q] for q in order], dtype=np.int32),
            "errors": np.array([errors[q] for q in order]), "global_mean": tot / cnt}
