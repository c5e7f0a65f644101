"""Plain data containers either side of the hot path (the subset of reference
pymht/utils/classDefinitions.py the path touches: MeasurementList :560-595, AisMessageList :597,
SimTargetCartesian :86-140, ScanList/SimList)."""
import numpy as np

from ..models import pv


class MeasurementList:
    """One radar scan: `.time` and `.measurements` (M,2)."""

    def __init__(self, time, measurements=None):
        self.time = time
        self.measurements = measurements if measurements is not None else []

    def __eq__(self, other):
        return self.time == other.time and np.array_equal(self.measurements, other.measurements)

    def __len__(self):
        return len(self.measurements)

    def filterUnused(self, unused_measurement_indices):
        return MeasurementList(self.time, np.asarray(self.measurements)[np.where(unused_measurement_indices)])

    def getMeasurements(self):
        return self.measurements


class AisMessageList(list):
    """AIS fusion is outside the accelerated path: only the empty list is accepted by Tracker."""

    def filterUnused(self, usedMmsiSet):
        return [m for m in self if m.mmsi not in usedMmsiSet]


class ScanList(list):
    pass


class SimList(list):
    pass


class SimTargetCartesian:
    """Ground-truth target for scenario generation (state [x,y,vx,vy], constant velocity + noise)."""

    def __init__(self, state, time, P_d, sigma_Q, **kwargs):
        self.state = np.array(state, dtype=np.double)
        self.time = time
        self.P_d = P_d
        self.sigma_Q = sigma_Q
        self.mmsi = kwargs.get("mmsi")
        self.model = pv

    def cartesianState(self):
        return self.state

    def inRange(self, p0, rRange):
        return np.linalg.norm(self.state[0:2] - p0) <= rRange
