"""Transform helpers with the surface of the reference's ``TransformHelper``
(transform_helper.py:14-259).  The 3x3 / 4x4 algebra of the motion predictor stays
on the host in numpy exactly as in the reference (it is a few dozen flops per
frame); projections go through the CUDA library (``cv_compat.projectPoints``)."""
from math import atan2, cos, sin, sqrt
from typing import Dict, List, Tuple

import numpy as np

from .. import cv_compat as agt_cv


class TransformHelper:
    def __init__(self, logger, mtx, dist):
        self.logger = logger
        self.mtx = mtx
        self.dist = dist

    # -- containers / geometry (transform_helper.py:29-96) ----------------------------
    @staticmethod
    def add_values_in_dict(sample_dict: Dict, key: int, list_of_values: List[object]) -> Dict:
        sample_dict.setdefault(key, []).extend(list_of_values)
        return sample_dict

    @staticmethod
    def get_initial_pts(tagsize: float) -> np.ndarray:
        """Corner order (-,-), (-,+), (+,+), (+,-): this order IS the corner indexing contract."""
        h = tagsize / 2.0
        return np.array([[-h, -h, 0.0], [-h, h, 0.0], [h, h, 0.0], [h, -h, 0.0]])

    @staticmethod
    def transform_marker_corners(object_pts: np.ndarray, transformation: Tuple[np.ndarray, np.ndarray]) -> np.ndarray:
        rvec, tvec = transformation
        if rvec.size == 0 or tvec.size == 0:
            raise ValueError('The transform rotation or translation: {} entered is empty'.format(transformation))
        rmat = agt_cv.Rodrigues(rvec)[0]
        return object_pts @ rmat.T + tvec.reshape(-1, 3)

    # -- reprojection gate (transform_helper.py:98-121) ---------------------------------
    def get_reprojection_error(self, obj_points, img_points, transformation) -> float:
        proj, _ = agt_cv.projectPoints(obj_points, transformation[0], transformation[1], self.mtx, self.dist)
        proj = proj.reshape(-1, 2)
        total = sum(np.linalg.norm(img_points[i] - proj[i]) for i in range(len(proj)))
        return total / len(proj)

    # -- 4x4 helpers (transform_helper.py:123-164) ---------------------------------------
    def get_extrinsic_matrix(self, rmat: np.ndarray, tvec: np.ndarray) -> np.ndarray:
        try:
            top = np.hstack((rmat, tvec))
            return np.vstack((top, np.array([0, 0, 0, 1])))
        except ValueError as err:
            raise ValueError('The rotation matrix or translation vector entered are not in the right format '
                             '(3x3 matrix and 3x1 vector) or are zero.') from err

    @staticmethod
    def get_rmat_tvec(extrinsic_mat: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        try:
            rot = extrinsic_mat[0:3, 0:3]
            tvec = np.array(extrinsic_mat[0:3, 3], dtype=np.float32).reshape(3, -1)   # float32 as in the reference
        except ValueError as err:
            raise ValueError('The extrinsic matrix entered is not a 4x4 matrix or is zero.') from err
        return rot, tvec

    # -- relative motion (transform_helper.py:166-212) -------------------------------------
    @staticmethod
    def get_relative_trans(rot_mat, tvec1, tvec0):
        try:
            return rot_mat.T @ (tvec0 - tvec1)
        except ValueError as err:
            raise ValueError('The vectors entered are either not the same size or zero.') from err

    @staticmethod
    def get_relative_rot(rmat0, rmat1):
        try:
            return rmat1.T @ rmat0
        except ValueError as err:
            raise ValueError('The matrices entered are either not the same size or zero.') from err

    # -- Euler angles, R = Rz Ry Rx (transform_helper.py:214-259) ---------------------------
    @staticmethod
    def euler_angles_to_rotation_matrix(theta) -> np.ndarray:
        cx, sx = cos(theta[0]), sin(theta[0])
        cy, sy = cos(theta[1]), sin(theta[1])
        cz, sz = cos(theta[2]), sin(theta[2])
        r_x = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
        r_y = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
        r_z = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
        return np.dot(r_z, np.dot(r_y, r_x))

    @staticmethod
    def rotation_matrix_to_euler_angles(rmat: np.ndarray) -> np.ndarray:
        s_y = sqrt(rmat[0, 0] * rmat[0, 0] + rmat[1, 0] * rmat[1, 0])
        if s_y >= 1e-6:
            return np.array([atan2(rmat[2, 1], rmat[2, 2]), atan2(-rmat[2, 0], s_y), atan2(rmat[1, 0], rmat[0, 0])])
        return np.array([atan2(-rmat[1, 2], rmat[1, 1]), atan2(-rmat[2, 0], s_y), 0])
