"""RenderArgs for a fixed camera looking at the effect grid (shared by bench.py and tests).
The frustum is built the way sf::Frustum's constructor does (src/sf/Frustum.h:16-40) with IEEE 1/sqrt."""
import numpy as np

from spear_b200 import abi

CAMERA = (16.0, 20.0, -30.0)


def look_at_perspective(eye, target, fovy=1.0, aspect=16 / 9, near=0.1, far=500.0):
    eye, target = np.asarray(eye, np.float64), np.asarray(target, np.float64)
    f = target - eye
    f /= np.linalg.norm(f)
    r = np.cross(f, [0, 1, 0])
    r /= np.linalg.norm(r)
    u = np.cross(r, f)
    view = np.eye(4)
    view[0, :3], view[1, :3], view[2, :3] = r, u, -f
    view[:3, 3] = -view[:3, :3] @ eye
    t = 1.0 / np.tan(fovy / 2)
    proj = np.zeros((4, 4))
    proj[0, 0], proj[1, 1] = t / aspect, t
    proj[2, 2], proj[2, 3], proj[3, 2] = (far + near) / (near - far), 2 * far * near / (near - far), -1
    return proj, proj @ view


def frustum_from_matrix(m16, clip_near_w=0.0):
    """sf::Frustum(const Mat44&, float) restated with float32 ops and 1/sqrt."""
    f32 = np.float32
    m = np.asarray(m16, dtype=f32).reshape(4, 4)  # m[col][row]
    r = [m[:, k].copy() for k in range(4)]        # r[k][j] = cols[j].v[k]
    fr = abi.Frustum()

    def fill(dst, planes):
        for pl, p in enumerate(planes):
            rcp = f32(1.0) / np.sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2], dtype=f32)
            for comp in range(4):
                dst[comp][pl] = float(f32(p[comp] * rcp))
    fill(fr.side, [r[3] - r[0], r[3] + r[0], r[3] - r[1], r[3] + r[1]])
    a = r[3] * f32(-clip_near_w) + r[2]
    b = r[3] - r[2]
    fill(fr.caps, [a, b, a, b])
    return fr


def make_render_args(camera=CAMERA, target=(14.0, 0.0, 14.0)):
    proj, w2c = look_at_perspective(camera, target)
    ra = abi.RenderArgs()
    ra.cameraPosition = abi.Vec3(*[float(x) for x in camera])
    for c in range(4):
        for r in range(4):
            ra.viewToClip.v[c * 4 + r] = float(np.float32(proj[r, c]))
            ra.worldToClip.v[c * 4 + r] = float(np.float32(w2c[r, c]))
    ra.frustum = frustum_from_matrix(list(ra.worldToClip.v))
    ra.visFogWorldMad = abi.Vec4(0.1, 0.2, 0.3, 0.4)
    return ra


def area_test_frusta():
    """Four frusta for the area-query tests (row N5): the bench camera, a top-down view, a near-horizontal view
    from inside the scene, and a narrow view from far away."""
    views = [((16.0, 20.0, -30.0), (14.0, 0.0, 14.0)), ((0.5, 80.0, 0.25), (0.0, 0.0, 0.0)),
             ((-20.0, 2.0, 5.0), (30.0, 1.0, -10.0)), ((150.0, 40.0, 150.0), (0.0, 0.0, 0.0))]
    return [make_render_args(camera=c, target=t).frustum for c, t in views]
