"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement (torch fp32 on the host, numpy for the integer graph work) of the
GeoSSL-DDM pretraining hot path of chao1224/GeoSSL.  It is the checker for the CUDA
product in ``geossl_b200/``; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The product
never routes through this package and has no CPU fallback.

Parity status
-------------
* First-party reference functions (SchNet, PaiNN, NCSN_version_03, do_DDM/perturb) are
  PINNED: ``tests/golden/make_golden.py`` imports the unmodified reference modules from
  /root/reference under the shims in ``oracle/shims`` and stores their outputs; the
  restatement is checked against those fixtures in ``tests/test_oracle_golden.py``.
* Third-party arithmetic (torch_cluster.radius_graph, torch_scatter.scatter,
  PyG MessagePassing.propagate) is un-vendored and absent from /root/reference and from
  this image: **parity unpinned** for those three.  ``oracle/radius.py`` restates the
  published torch_cluster CUDA semantics (index-order scan, 33-candidate truncation,
  strict ``d2 < r*r``, self removed, target-major / source-ascending output).
"""
