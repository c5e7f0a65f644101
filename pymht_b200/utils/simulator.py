"""Seeded synthetic scenario generator with the reference simulator's call names
(pymht/utils/simulator.py:15-110): uniformly placed CV targets in a disc, detections with
probability P_d plus Poisson clutter of density lambda_phi.  Own implementation (vectorised, its
own RandomState); used by bench.py and the tests as the input generator -- not part of the path."""
import numpy as np

from ..models import pv
from .classDefinitions import MeasurementList, ScanList, SimList, SimTargetCartesian

_rng = np.random.RandomState(0)


def seed_simulator(seed):
    global _rng
    _rng = np.random.RandomState(seed)


def generateInitialTargets(numOfTargets, centerPosition, radarRange, P_d, sigma_Q, initialTime=0.0, **kwargs):
    speeds = np.array([1, 10, 12, 15, 28, 35], dtype=np.float64) * 0.5
    out = []
    for _ in range(numOfTargets):
        a = _rng.uniform(0, 2 * np.pi)
        d = _rng.uniform(0, radarRange * 0.8)
        h = _rng.uniform(0, 2 * np.pi)
        v = _rng.choice(speeds)
        state = [centerPosition[0] + d * np.cos(a), centerPosition[1] + d * np.sin(a), v * np.cos(h), v * np.sin(h)]
        out.append(SimTargetCartesian(np.array(state, dtype=np.float32), initialTime, P_d, sigma_Q))
    return out


def simulateTargets(initialTargets, simTime, timeStep, model=pv, **kwargs):
    simList = SimList()
    simList.append(list(initialTargets))
    A = np.asarray(model.Phi(timeStep), dtype=np.float64)
    # white-acceleration process noise on velocity, integrated into position
    for _ in range(int(np.ceil(simTime / timeStep))):
        nxt = []
        for tgt in simList[-1]:
            w = _rng.normal(scale=tgt.sigma_Q * np.sqrt(timeStep), size=2)
            state = A.dot(tgt.state)
            state[0:2] += 0.5 * timeStep * w
            state[2:4] += w
            nxt.append(SimTargetCartesian(state, tgt.time + timeStep, tgt.P_d, tgt.sigma_Q))
        simList.append(nxt)
    return simList


def simulateScans(simList, radarPeriod, H, R, lambda_phi=0, rRange=None, p0=None, **kwargs):
    skip_first = kwargs.get("preInitialized", False)
    sigma = float(np.sqrt(np.asarray(R, dtype=np.float64)[0, 0]))
    H = np.asarray(H, dtype=np.float64)
    scans = ScanList()
    for k, targets in enumerate(simList):
        if skip_first and k == 0:
            continue
        pts = []
        for tgt in targets:
            if _rng.uniform() <= kwargs.get("P_d", tgt.P_d) and (rRange is None or tgt.inRange(p0, rRange)):
                pts.append(H.dot(tgt.state) + _rng.normal(scale=sigma, size=2))
        if rRange is not None and p0 is not None and kwargs.get("globalClutter", True) and lambda_phi > 0:
            n = _rng.poisson(lambda_phi * np.pi * rRange ** 2)
            r = rRange * np.sqrt(_rng.uniform(size=n))
            th = _rng.uniform(0, 2 * np.pi, size=n)
            pts.extend(np.stack([p0[0] + r * np.cos(th), p0[1] + r * np.sin(th)], axis=1))
        pts = np.array(pts, dtype=np.float32).reshape(-1, 2)
        if kwargs.get("shuffle", True):
            _rng.shuffle(pts)
        scans.append(MeasurementList(targets[0].time, pts))
    return scans
