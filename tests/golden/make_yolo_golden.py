#!/usr/bin/env python
"""Generate tests/golden/yolo_v1.npz: two small synthetic network outputs (320x320 tensor, 1 and 3 classes) and the detections
the COMPILED REFERENCE post-processing (oracle/_ref/libref_yolo.so, built from detectors/yolo3.cpp:141-356) makes of them.
Run in the authoring container after `make -C oracle ref`."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oraclelib                      # noqa: E402
import test_oracle_yolo as T          # noqa: E402


def main():
    ref = C.CDLL(os.path.join(oraclelib.ODIR, "_ref", "libref_yolo.so"))
    rng = np.random.default_rng(20261017)
    out = {}
    for tag, nc, nobj, obj, nms, ih, iw in (("a", 1, 25, 0.5, 0.45, 720, 1280), ("b", 3, 60, 0.3, 0.4, 1080, 1920)):
        outs = T.synth_outputs(rng, 320, 320, nc, nobj)
        det = T.run(ref, "ref_yolo_post", outs, obj, nms, 320, 320, ih, iw, nc)
        assert len(det) > 5
        for k in range(3):
            out["%s_out%d" % (tag, k)] = outs[k]
        out[tag + "_det"] = det
        out[tag + "_cfg"] = np.array([320, 320, ih, iw, nc], np.int32)
        out[tag + "_thr"] = np.array([obj, nms], np.float32)
    np.savez_compressed(os.path.join(HERE, "yolo_v1.npz"), **out)
    print("wrote yolo_v1.npz:", {k: (v.shape, str(v.dtype)) for k, v in out.items() if k.endswith("det")})


if __name__ == "__main__":
    main()
