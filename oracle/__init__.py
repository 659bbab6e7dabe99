"""CPU oracle for the Polychase analyze/track/refine hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import it, and only as the checker (or as the timed CPU baseline), never as the
thing shipped.  The product path is the CUDA library behind
``include/polychase_b200.h`` and fails loudly when that library is missing.

Parity pinning: the reference has no tests and cannot be
compiled here (Eigen, oneTBB, Embree, OpenCV C++ headers, spdlog and sqlite3.h
are absent), so the oracle is pinned as follows:

* OpenCV-resident stages (RGB->gray, pyrDown, Scharr, cornerMinEigenVal,
  dilate/threshold, calcOpticalFlowPyrLK) are checked against ``cv2`` 4.13.0 --
  OpenCV itself, the third-party dependency the reference calls
  (reference pin: vcpkg tag 2025.06.13, opencv4 4.11.x).
* The LM machinery is checked against the reference's one recorded known-answer
  (cpp/examples/levmarq_ill_conditioned_float32_issue.cpp:16-63).
* Everything else Polychase-owned (GFTT grid logic, status filter, DB blobs, ray
  cast, PnP problem, BA residual/Jacobian) is a restatement with no reference
  golden vectors available: **parity unpinned** for those parts, beyond
  analytic-vs-numeric Jacobian checks and ground-truth recovery on synthetic
  scenes.
"""
