"""oracle/ -- TEST INFRASTRUCTURE ONLY (CPU restatement of the Kimera-VIO stereo front-end hot path).

This package restates, step by step, the reference front-end:
the reference's own glue logic is re-written from the cited file:line ranges, and every OpenCV
call the reference makes is made here through the `cv2` 4.13 wheel (same C++ implementation the
reference links).  The OpenGV RANSAC half (not available anywhere in this image) is restated in
numpy from the published algorithm -- see oracle/ransac.py header: that part is "parity pinned by
the reference's own synthetic-scene tests only" (tests/testTracker.cpp), not by a reference binary.

Modules: frontend.py (rows a2-a15 and the stereo front-end FSM), rig.py (a1, radtan and equidistant), ransac.py (the
OpenGV / GTSAM half), mono.py (MonoVisionImuFrontend), rgbd.py (DepthFrame / RgbdFrame functions and the
RgbdVisionImuFrontend flow), mesher.py (Mesher 2-D Delaunay), fisheye.py and maps.py (CPU twins of the device camera-model
formulas, pinned against cv2), rng_check.cpp / ref_rng.cpp (libstdc++ / glibc random streams).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this package, and only as the checker / the timed CPU baseline.  The product path
(kimera_vio_b200 -> libkvfe.so) never imports or links anything from here and fails loudly when the
CUDA library is missing.
"""
