"""One-off converter: FLAME `generic_model.pkl` -> `head_detector_b200/assets/flame_generic.npz`.

The reference loads the pickle at run time (head_detector/flame.py:18-24, needs chumpy +
scipy); the B200 build wants plain fp32 arrays it can upload once.  The arrays written are
exactly the buffers `FLAMELayer.__init__` registers (head_detector/flame.py:43-95), cast to
fp32 the same way (`to_np(...)` then `to_tensor(dtype=float32)`).

Usage:  python tools/convert_flame_pkl.py [/path/to/generic_model.pkl] [out.npz]
"""
import pickle
import sys
import types
import os

import numpy as np


def _install_chumpy_stub():
    """The pickle references `chumpy.ch.Ch` (only for `shapedirs`); its ndarray lives in state['x']."""
    if "chumpy" in sys.modules:
        return

    class Ch:
        def __setstate__(self, state):
            self.__dict__.update(state)

    pkg, sub = types.ModuleType("chumpy"), types.ModuleType("chumpy.ch")
    sub.Ch = Ch
    pkg.ch = sub
    sys.modules["chumpy"], sys.modules["chumpy.ch"] = pkg, sub


def convert(pkl_path: str, out_path: str) -> None:
    _install_chumpy_stub()
    with open(pkl_path, "rb") as f:
        m = pickle.load(f, encoding="latin1")
    shapedirs = m["shapedirs"]
    shapedirs = np.asarray(shapedirs.x if hasattr(shapedirs, "x") else shapedirs)
    posedirs = np.asarray(m["posedirs"])
    parents = np.asarray(m["kintree_table"][0]).astype(np.int64)
    parents[0] = -1
    arrays = dict(
        v_template=np.asarray(m["v_template"], dtype=np.float32),
        shapedirs=shapedirs.astype(np.float32),
        posedirs=np.reshape(posedirs, [-1, posedirs.shape[-1]]).T.astype(np.float32).copy(),
        J_regressor=np.asarray(m["J_regressor"].todense(), dtype=np.float32),
        lbs_weights=np.asarray(m["weights"], dtype=np.float32),
        parents=parents,
        faces=np.asarray(m["f"]).astype(np.int32),
    )
    assert arrays["v_template"].shape == (5023, 3) and arrays["shapedirs"].shape == (5023, 3, 400)
    assert arrays["posedirs"].shape == (36, 15069) and arrays["J_regressor"].shape == (5, 5023)
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    np.savez_compressed(out_path, **arrays)


if __name__ == "__main__":
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/head_detector/generic_model.pkl"
    dst = sys.argv[2] if len(sys.argv) > 2 else os.path.join(here, "head_detector_b200", "assets", "flame_generic.npz")
    convert(src, dst)
    print("wrote", dst, os.path.getsize(dst), "bytes")
