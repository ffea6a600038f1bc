"""Stand-in for tf: TransformerROS.fromTranslationRotation (translation + quaternion -> 4x4)."""
import numpy as np


def quaternion_matrix(q):
    q = np.array(q, dtype=np.float64)
    nq = float(np.dot(q, q))
    if nq < np.finfo(np.float64).eps * 4.0:
        return np.identity(4)
    q = q * np.sqrt(2.0 / nq)
    o = np.outer(q, q)                       # q = [x, y, z, w]
    return np.array([[1.0 - o[1, 1] - o[2, 2], o[0, 1] - o[2, 3], o[0, 2] + o[1, 3], 0.0],
                     [o[0, 1] + o[2, 3], 1.0 - o[0, 0] - o[2, 2], o[1, 2] - o[0, 3], 0.0],
                     [o[0, 2] - o[1, 3], o[1, 2] + o[0, 3], 1.0 - o[0, 0] - o[1, 1], 0.0],
                     [0.0, 0.0, 0.0, 1.0]])


class TransformerROS:
    def fromTranslationRotation(self, translation, rotation):
        m = quaternion_matrix(rotation)
        m[:3, 3] = translation
        return m
