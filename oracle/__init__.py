"""ORACLE — test infrastructure only.

CPU restatements of the reference's algorithms for the hot path.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package, and only as the checker / the timed CPU arm — never from
``benchpush_b200`` (the product), which must fail loudly when its CUDA library is missing.

Parity status:
  * physics (``chipmunk_oracle.cpp``): PARITY UNPINNED — restates third-party pymunk 6.5.1 /
    Chipmunk2D 7.0.3, absent from the reference tree and from this image; the reference has no
    golden vectors for it.
  * rasters (``raster_oracle.py``): cv2 / scipy primitives are called directly (importable,
    bit-exact); ``skimage.draw.polygon`` is restated (skimage absent) — unpinned for that rule.
  * spfa (``spfa_oracle.cpp``): PINNED against the reference source compiled in place
    (``oracle/_ref``) and the committed ``tests/golden/spfa_*.npz``.
"""
