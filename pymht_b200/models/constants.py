"""Model constants (reference pymht/models/constants.py:2-10)."""
import numpy as np

defaultType = np.float32     # the reference keeps every model matrix in float32
nDimState = 4
sigmaR_RADAR_tracker = 2.5   # measurement std used by the filter
sigmaR_RADAR_true = 2.5
sigmaQ_tracker = 1.0         # process-noise scale used by the filter
sigmaQ_true = 1.0
