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

Parity status: the reference's own tests hold no golden vectors for this path
(SURVEY.md §4, §8c); parity is pinned to OpenCV itself via cv2 4.13.0.
"""
