"""Result serialisation: skeleton records -> the reference's result JSON.

Mirrors what generate_3d_point_pairs writes in run_inference mode (exps/stage3_root2/test.py:32-34,145-152 and
save_result, exps/stage3_root2/test_util.py:146-158): {"model_pattern": NAME, "3d_pairs": [{"pred_2d", "pred_3d",
"root_d", "image_path", "gt_3d": [], "gt_2d": []}, ...]}, one entry per image with at least one person.  The writing
is done by libsmap_b200.so (smapb_json_*), byte-identical to json.dump and ~50x faster than the Python encoder, which
would otherwise cap the pipeline below the GPU's frame rate.
"""
import ctypes
import os

import numpy as np

from . import _lib
from ._lib import SmapB200Error


def result_file_name(output_dir, test_mode="run_inference", data_mode="test", suffix="", dir_name="stage3_root2"):
    """'{dir}_{TEST_MODE}_{DATA_MODE}_{JSON_SUFFIX_NAME}.json' (exps/stage3_root2/test.py:147-149)."""
    return os.path.join(output_dir, "{}_{}_{}_{}.json".format(dir_name, test_mode, data_mode, suffix))


class ResultWriter:
    """with ResultWriter(path, cfg.DATASET.NAME) as w: w.append(records, image_paths) for every batch."""

    def __init__(self, path, model_pattern):
        self.lib = _lib.load()
        self._w = ctypes.c_void_p()
        rc = self.lib.smapb_json_open(ctypes.byref(self._w), os.fsencode(path), str(model_pattern).encode("utf-8"))
        if rc != 0:
            raise SmapB200Error("smapb_json_open(%s) failed (%d)" % (path, rc))

    def append(self, records, image_paths):
        """records: host records of one batch - a uint8 array/tensor [B, RECORD_BYTES] or a structured array with
        engine.RECORD_DTYPE; image_paths: B strings."""
        if hasattr(records, "numpy"):
            records = records.numpy()
        a = np.ascontiguousarray(records)
        B = a.shape[0]
        if len(image_paths) != B:
            raise ValueError("%d records but %d image paths" % (B, len(image_paths)))
        paths = (ctypes.c_char_p * max(1, B))(*[str(p).encode("utf-8") for p in image_paths])
        rc = self.lib.smapb_json_append(self._w, a.ctypes.data_as(ctypes.c_void_p), B, paths)
        if rc != 0:
            raise SmapB200Error("smapb_json_append failed (%d)" % rc)

    def close(self):
        if self._w:
            rc = self.lib.smapb_json_close(self._w)
            self._w = ctypes.c_void_p()
            if rc != 0:
                raise SmapB200Error("smapb_json_close failed (%d)" % rc)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
