"""Drop-in for the reference's ``Geom3D`` package (models only -- the part on the DDM hot path)."""
