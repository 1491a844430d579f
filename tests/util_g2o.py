"""Write a measurement set back to .g2o text (tests only).  The information matrices are chosen
so that the reference's parser formulas (src/DPGO_utils.cpp:174-176, :223-230) return exactly
the given kappa / tau: translation block tau*I, rotation block 2*kappa*I (3-D) or kappa (2-D)."""
import numpy as np


def rot_to_quat(R):
    """(x, y, z, w) of a rotation matrix (Shepperd)."""
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        w, x, y, z = 0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = np.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        w, x, y, z = (R[2, 1] - R[1, 2]) / s, 0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s
    elif R[1, 1] > R[2, 2]:
        s = np.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        w, x, y, z = (R[0, 2] - R[2, 0]) / s, (R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s
    else:
        s = np.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        w, x, y, z = (R[1, 0] - R[0, 1]) / s, (R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s
    return x, y, z, w


def write_g2o(path, d, p1, p2, R, t, kappa, tau):
    g = lambda v: repr(float(v))
    with open(path, "w") as fh:
        for k in range(len(p1)):
            if d == 3:
                x, y, z, w = rot_to_quat(R[k])
                info = np.zeros((6, 6))
                info[:3, :3] = tau[k] * np.eye(3)
                info[3:, 3:] = 2 * kappa[k] * np.eye(3)
                up = [info[i, j] for i in range(6) for j in range(i, 6)]
                fh.write("EDGE_SE3:QUAT %d %d %s %s\n" % (
                    p1[k], p2[k], " ".join(g(v) for v in list(t[k]) + [x, y, z, w]), " ".join(g(v) for v in up)))
            else:
                th = np.arctan2(R[k][1, 0], R[k][0, 0])
                fh.write("EDGE_SE2 %d %d %s %s %s %s 0.0 0.0 %s 0.0 %s\n" % (
                    p1[k], p2[k], g(t[k][0]), g(t[k][1]), g(th), g(tau[k]), g(tau[k]), g(kappa[k])))
