"""Minimal stand-in for ``ase.data``: only ``atomic_masses`` (119 float64 entries) is read
by the reference (Geom3D/models/schnet.py:47) and only its shape/dtype matter on the hot path
(the buffer is used under ``dipole=True`` only)."""
import numpy as np

from geossl_b200.atomic_data import ATOMIC_MASSES

atomic_masses = np.asarray(ATOMIC_MASSES, dtype=np.float64)
