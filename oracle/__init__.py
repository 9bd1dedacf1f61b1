"""CPU oracle for the BodySLAM depth->3D hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product
(``bodyslam_b200``) never does; it fails loudly without its CUDA library.

* ``oracle.o3d``  -- ctypes wrapper over ``libo3d_oracle.so`` (C + OpenMP restatement of the
  Open3D legacy-pipeline arithmetic the reference calls from ``N/3DM/tsdf.py`` /
  ``N/3DM/slam_utils.py``; parity vs. real Open3D is UNPINNED, see ``o3d_oracle.c``).
* ``oracle.mdem`` -- NumPy restatement of the MDEM post-processing
  (``R/examples/depth_estimation/depth_map_scaling.py:12-45``), pinned by the reference's
  golden pair (``tests/golden/colorize_golden.npz``).
"""
from . import mdem, o3d  # noqa: F401
