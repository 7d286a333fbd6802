"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the scannertools per-frame analysis hot path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
leg may import anything from this package.  The product (`scannertools_b200/`) never does.

Two layers:
  * `oracle.cv2_ops`  -- drives the SAME OpenCV functions with the SAME arguments as the
    reference's C++ wrappers (OpenCV is the un-vendored third-party dependency where the
    arithmetic lives; cv2 4.13.0 is the copy available in this image).
  * `oracle.restate`  -- ctypes loader of `oracle/restate.c`, an OpenCV-free plain-C
    restatement, pinned against goldens produced by `cv2_ops`
    (tests/golden/, generator script committed beside them).

Parity status, precisely:
  * The reference's own tests hold NO golden vectors for this path, and its
    C++ wrappers cannot be compiled or run here (they need the Scanner engine and OpenCV C++).
  * What IS pinned: (1) the arithmetic, against outputs of OpenCV itself (the un-vendored
    dependency that implements it) produced here through cv2 4.13.0 with the wrappers' exact call
    arguments -- tests/golden/*.npz; (2) the one Python op on the path, ShotBoundaries, against
    outputs of the REFERENCE ITSELF: scannertools/shot_detection.py imported from /root/reference
    behind a scannerpy stub (tests/golden/ref_import.py) -> tests/golden/shot_reference.npz.
  * What is not pinned by anything the reference ships: the C++ wrappers' behaviour beyond what
    their source shows (argument order, output layout), e.g. OpenCV-version drift.
"""
