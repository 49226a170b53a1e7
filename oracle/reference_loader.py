"""Imports the UNMODIFIED reference modules from /root/reference under oracle/shims.
Only usable where /root/reference exists (the build container); used by
tests/golden/make_golden.py and by the optional cross-checks in tests/test_oracle_golden.py.
TEST INFRASTRUCTURE."""
import os
import sys
import warnings

REFERENCE_ROOT = os.environ.get("GEOSSL_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "Geom3D", "models"))


def load():
    """Returns (SchNet, PaiNN, NCSN_version_03) classes of the reference."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    here = os.path.dirname(os.path.abspath(__file__))
    repo = os.path.dirname(here)
    for p in (repo, os.path.join(here, "shims"), REFERENCE_ROOT, os.path.join(REFERENCE_ROOT, "examples")):
        if p not in sys.path:
            sys.path.insert(0, p)
    # the product package also ships a drop-in `Geom3D`; make sure the reference's one wins here
    for name in [m for m in sys.modules if m == "Geom3D" or m.startswith("Geom3D.")]:
        del sys.modules[name]
    sys.path.remove(REFERENCE_ROOT)
    sys.path.insert(0, REFERENCE_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from Geom3D.models import SchNet, PaiNN          # noqa: E402
        from NCSN import NCSN_version_03                 # noqa: E402
    return SchNet, PaiNN, NCSN_version_03
