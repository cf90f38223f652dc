#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY.  Writes, into a scratch directory OUTSIDE the repository, an include file with the text of the
YOLO3 post-processing of the reference exactly as it stands in detectors/yolo3.cpp: the two POD typedefs (:96-109) and the
four static functions decode_netout, correct_yolo_boxes, sort, do_nms (:141-356).  The file cannot be compiled as a whole
(TensorFlow, OpenCV, Windows), so oracle/Makefile compiles this excerpt where it lies, together with oracle/capi/yolo_capi.cpp,
into oracle/_ref/libref_yolo.so.  Nothing of the reference is copied into the repository.

usage: extract_yolo.py /root/reference/detectors/yolo3.cpp OUT.inc"""
import re
import sys


def main():
    src = open(sys.argv[1], "rb").read().decode("latin-1")
    t0 = src.index("typedef struct {\r\n\tfloat x, y, u, w;") if "typedef struct {\r\n\tfloat x, y, u, w;" in src else src.index("typedef struct {\n\tfloat x, y, u, w;")
    t1 = src.index("} detection_t;") + len("} detection_t;")
    f0 = src.index("static void decode_netout(")
    f1 = src.index("std::unique_ptr<tensorflow::Session> session;")
    out = "// generated from detectors/yolo3.cpp by oracle/tools/extract_yolo.py -- do not commit\n" + src[t0:t1] + "\n\n" + src[f0:f1]
    open(sys.argv[2], "wb").write(out.encode("latin-1"))
    names = re.findall(r"static void (\w+)\(", src[f0:f1])
    assert names == ["decode_netout", "correct_yolo_boxes", "sort", "do_nms"], names


if __name__ == "__main__":
    main()
