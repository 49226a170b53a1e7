"""Recipe: populate the git-ignored ``oracle/_ref/`` with the reference's OWN, UNMODIFIED module files.

    python oracle/make_ref.py            # run in the build container (needs /root/reference)

TEST INFRASTRUCTURE.  The reference is pure Python (no build system), so "building" it is
placing the five files its DDM path needs where they can be imported on the GPU box, which has no
/root/reference: ``oracle/_ref/`` is listed in .gitignore (the sources never enter this repo's history) but not in
.gpurunignore, so it travels with the snapshot exactly like the built ``libgeossl_b200.so``.
``oracle/reference_loader.py`` imports them under the import shims in ``oracle/shims`` (ase / torch_geometric /
torch_scatter / torch_cluster are absent from the image); ``bench.py --impl reference`` and the ``cpu_baseline`` leg
time these modules (``kind: "reference"``) and fall back to the oracle port only when ``oracle/_ref`` is absent.
``__graft_entry__.build()`` runs this recipe whenever /root/reference is present.
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("GEOSSL_REFERENCE_ROOT", "/root/reference")
DEST = os.path.join(HERE, "_ref")
FILES = ("Geom3D/models/__init__.py", "Geom3D/models/schnet.py", "Geom3D/models/painn.py",
         "Geom3D/models/painn_utils.py", "examples/NCSN.py")


def make(verbose=True):
    if not os.path.isdir(os.path.join(REF_SRC, "Geom3D", "models")):
        if verbose:
            print(f"oracle/make_ref.py: no reference tree at {REF_SRC}; oracle/_ref left as it is")
        return False
    for rel in FILES:
        src, dst = os.path.join(REF_SRC, rel), os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.exists(dst) and filecmp.cmp(src, dst, shallow=False)):
            shutil.copyfile(src, dst)
    if verbose:
        print(f"oracle/_ref: {len(FILES)} unmodified reference files from {REF_SRC}")
    return True


if __name__ == "__main__":
    sys.exit(0 if make() else 1)
