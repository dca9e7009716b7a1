"""Drop-in counterpart of the reference package ``aprilgroup_pose_estimation``
(aprilgroup_tracking/aprilgroup_pose_estimation): same class and method names,
same attributes, with the OpenCV arithmetic of the hot path executed by libagt.so."""
from .detect_pose import PoseDetector  # noqa: F401
from .transform_helper import TransformHelper  # noqa: F401
