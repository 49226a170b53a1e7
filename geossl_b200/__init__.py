"""geossl_b200 -- B200-native (sm_100a) GeoSSL-DDM pretraining hot path.

Drop-in for the reference's ``Geom3D.models`` (SchNet / PaiNN), ``NCSN.NCSN_version_03`` and
``pretrain_GeoSSL.do_DDM``; all numerics run in the hand-written CUDA library
``libgeossl_b200.so`` (C ABI declared in include/geossl_b200.h).  There is no CPU fallback."""
__version__ = "0.1.0"
