"""Import stub so that `import head_detector` (the reference package) works in this
container without building its Cython rasteriser, which is not on the hot path
(reference: head_detector/Sim3DR/Sim3DR.py:6). TEST INFRASTRUCTURE ONLY."""


def _unavailable(*a, **k):
    raise RuntimeError("Sim3DR_Cython stub: the CPU rasteriser is out of scope for this build")


get_normal = rasterize = _unavailable
