"""Constant-velocity (position/velocity) model, state [x, y, vx, vy]
(same public names as reference pymht/models/pv.py:7-34; float32 like the reference)."""
import numpy as np

from .constants import *  # noqa: F401,F403
from .constants import defaultType, sigmaQ_tracker, sigmaR_RADAR_tracker

C_RADAR = np.zeros((2, 4), dtype=defaultType)
C_RADAR[0, 0] = C_RADAR[1, 1] = 1.0
H_radar = C_RADAR

_p0 = 2.5 ** 2
P0 = np.diag([_p0, _p0, 0.3 * _p0, 0.3 * _p0]).astype(defaultType)


def Phi(T):
    A = np.eye(4, dtype=defaultType)
    A[0, 2] = A[1, 3] = T
    return A


def Q(T, sigmaQ=sigmaQ_tracker):
    q = np.zeros((4, 4))
    q[0, 0] = q[1, 1] = T ** 4 / 4.0
    q[0, 2] = q[2, 0] = q[1, 3] = q[3, 1] = T ** 3 / 3.0
    q[2, 2] = q[3, 3] = T ** 2
    return q.astype(defaultType) * sigmaQ


def R_RADAR(sigmaR=sigmaR_RADAR_tracker):
    return (np.eye(2) * sigmaR ** 2).astype(defaultType)
