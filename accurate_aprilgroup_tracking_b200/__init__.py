"""B200-native AprilGroup tracking hot path (APE -> pyramidal LK -> dense pose refinement).

Host side is Python; all arithmetic of the path runs in hand-written CUDA for
sm_100a behind the C ABI declared in ``include/agt.h`` (``csrc/libagt.so``).
There is no CPU fallback: every compute entry point raises if the library or a
CUDA device is missing.
"""
__version__ = "0.1.0"
