"""pymht_b200: B200-native (sm_100a CUDA) implementation of pyMHT's per-scan hot path behind the
reference's Tracker.addMeasurementList / Target API.  See DESIGN.md.  No CPU fallback."""
__version__ = "0.1.0"
