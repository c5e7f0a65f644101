"""`Target`: one hypothesis-tree node with the reference's field names (pymht/pyTarget.py:16-40).

In the reference every node is a Python object.  Here the forest lives in HBM
(csrc/forest.cu) and a `Target` is either
  * a user-built initial target handed to Tracker.initiateTarget (plain host object), or
  * a lazily materialised VIEW of a device node: the selected leaves returned by
    Tracker.getTrackNodes(); `.parent` materialises the chain back to the initial node on first use.
"""
import copy

import numpy as np

activeTag = "Active"                 # reference pymht/utils/xmlDefinitions.py
preinitializedTag = "preinitialized"
outofrangeTag = "OutOfRange"
toolowscoreTag = "TooLowScore"
STATUS_TAGS = {0: activeTag, 1: outofrangeTag, 2: toolowscoreTag}


class Target:
    def __init__(self, time, scanNumber, x_0, P_0, ID=None, S_inv=None, **kwargs):
        x_0 = np.asarray(x_0)
        P_0 = np.asarray(P_0)
        assert (scanNumber is None) or (scanNumber == int(scanNumber))
        assert x_0.ndim == 1 and P_0.ndim == 2
        assert x_0.shape[0] == P_0.shape[0] == P_0.shape[1]
        self.isRoot = kwargs.get("isRoot", False)
        self.ID = ID
        self.time = time
        self.scanNumber = scanNumber
        self.x_0 = x_0
        self.P_0 = P_0
        self.S_inv = S_inv
        self.P_d = copy.copy(kwargs.get("P_d", 0.8))
        self._parent = kwargs.get("parent")
        self._parent_loader = kwargs.get("parent_loader")
        self.measurementNumber = kwargs.get("measurementNumber", 0)
        self.measurement = kwargs.get("measurement")
        self.cumulativeNLLR = copy.copy(kwargs.get("cumulativeNLLR", 0))
        self.trackHypotheses = None
        self.mmsi = kwargs.get("mmsi")
        self.status = kwargs.get("status", activeTag)
        assert 0 <= self.P_d <= 1

    # -- tree links ------------------------------------------------------------------------------
    @property
    def parent(self):
        if self._parent is None and self._parent_loader is not None:
            self._parent_loader(self)      # fills self._parent (and the whole chain) once
            self._parent_loader = None
        return self._parent

    @parent.setter
    def parent(self, value):
        self._parent = value
        self._parent_loader = None

    def getRoot(self):
        node = self
        while node is not None and not node.isRoot:
            node = node.parent
        return node

    def getInitial(self):
        node = self
        while node.parent is not None:
            node = node.parent
        return node

    def stepBack(self, stepsBack=1):
        node = self
        while stepsBack > 0 and node.parent is not None:
            node, stepsBack = node.parent, stepsBack - 1
        return node

    def height(self):
        n, node = 1, self
        while node.parent is not None:
            n, node = n + 1, node.parent
        return n

    def rootHeight(self):
        n, node = 0, self
        while node.parent is not None and not node.isRoot:
            n, node = n + 1, node.parent
        return n

    # -- values ----------------------------------------------------------------------------------
    def getScore(self):
        """cumulativeNLLR relative to the tree's current root (pyTarget.py:124-125)."""
        return self.cumulativeNLLR - self.getRoot().cumulativeNLLR

    def getPosition(self):
        return np.array(self.x_0[0:2])

    def getVelocity(self):
        return np.array(self.x_0[2:4])

    def isOutsideRange(self, position, range):
        return np.linalg.norm(self.x_0[0:2] - position) > range

    def __sub__(self, other):
        return self.x_0 - other.x_0

    def __repr__(self):
        return ("Target(ID=%s scan=%s meas=%s pos=(%.1f,%.1f) vel=(%.1f,%.1f) cNLLR=%.4f %s)" % (
            self.ID, self.scanNumber, self.measurementNumber, self.x_0[0], self.x_0[1], self.x_0[2], self.x_0[3],
            self.cumulativeNLLR, self.status))


def backtrackMeasurementNumbers(selectedNodes, steps=None):
    """measurementNumber history of each node (reference pymht/utils/helpFunctions.py:66-83)."""
    out = []
    for node in selectedNodes:
        hist, left = [], steps
        while node.parent is not None and (left is None or left > 0):
            hist.append(int(node.measurementNumber))
            node = node.parent
            if left is not None:
                left -= 1
        out.append(hist[::-1])
    return out
