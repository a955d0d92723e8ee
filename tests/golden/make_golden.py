#!/usr/bin/env python
"""Regenerates tests/golden/ from the read-only reference checkout.

The hot path's only reference artefacts are DATA files committed under
/root/reference/Common: inputs of the bundle adjustment
(correspondence files, intrinsics) and the outputs the reference itself wrote after
running Ceres on them (Camera_Transform.xml with 17 digits; point3d.txt and
Extrinsics/mat*.txt with 6 digits).  They cannot be read at test time on the GPU box
(/root/reference does not exist there), so this script copies them verbatim into
tests/golden/reference_common/ keeping their relative paths.  No reference SOURCE
file is copied.  Run:  python tests/golden/make_golden.py
"""
import os
import shutil
import sys

REF = os.environ.get("BA_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
FILES = [
    "Common/Correspondence/hongo/correspondence.txt",
    "Common/Correspondence/hongo/Camera_Transform.xml",
    "Common/Correspondence/hongo/point3d.txt",
    "Common/Correspondence/hongo/marker_geometry.txt",
    "Common/Correspondence/test2/correspondence_test.txt",
    "Common/Correspondence/test2/Camera_Transform.xml",
    "Common/Correspondence/test2/point3d.txt",
    "Common/Correspondence/test2/geometry_test.txt",
    "Common/Correspondence/two_cam_data.txt",
    "Common/Calibration/Extrinsics/mat0.txt",
    "Common/Calibration/Extrinsics/mat1.txt",
    "Common/Calibration/Extrinsics/mat2.txt",
    "Common/Calibration/Extrinsics/mat3.txt",
] + ["Common/Calibration/Intrinsics/%s.xml" % s for s in (
    "821312061029", "816612062327", "821212062536", "821212061326", "819612072493", "825312072048")]


def main():
    if not os.path.isdir(REF):
        sys.exit("reference checkout not found at %s" % REF)
    for rel in FILES:
        dst = os.path.join(HERE, "reference_common", rel[len("Common/"):])
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF, rel), dst)
        print("copied", rel)


if __name__ == "__main__":
    main()
