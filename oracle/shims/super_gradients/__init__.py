"""Stand-in for the third-party `super_gradients` package (`super_gradients>=3.7`,
yolo_head_training/requirements.txt:1; absent from this image, no network).  TEST INFRASTRUCTURE ONLY.

Purpose: let the UNMODIFIED reference files
    yolo_head_training/yolo_head/yolo_head_dfl_head.py   (YoloHeadsDFLHead)
    yolo_head_training/yolo_head/yolo_head_ndfl_heads.py (YoloHeadsNDFLHeads: DFL decode, anchors, channel rotation)
    yolo_head_training/yolo_head/yolo_heads.py           (VGGHeadDecodingModule: top-k)
be imported and run here, so that golden vectors for the head / decode / top-k stages come from the reference's own
code (oracle/make_golden.py) - exactly what oracle/shims/smplx does for the FLAME layer.

What is restated from the library's published definitions ([3P-MEM]): `_impl.py` -
ConvBNReLU, QARepVGGBlock, BaseDetectionModule, width_multiplier, batch_distance2bbox,
generate_anchors_for_grid_cell, HpmStruct, DetectionModulesFactory, the registry decorators.  Every other name the
reference files import from `super_gradients.*` (export helpers, loggers, interfaces, onnx ...) is fabricated on
demand as an inert placeholder: those code paths (export(), training) are never executed here.
"""
import importlib.abc
import importlib.machinery
import sys
import types

from . import _impl

_REAL = {
    "super_gradients.common.registry": ("register_detection_module", "register_model"),
    "super_gradients.common.registry.registry": ("register_detection_module", "register_model"),
    "super_gradients.modules": ("ConvBNReLU", "QARepVGGBlock"),
    "super_gradients.modules.base_modules": ("BaseDetectionModule",),
    "super_gradients.modules.utils": ("width_multiplier",),
    "super_gradients.common.factories.detection_modules_factory": ("DetectionModulesFactory",),
    "super_gradients.training.models.detection_models.pp_yolo_e.pp_yolo_head": ("generate_anchors_for_grid_cell",),
    "super_gradients.training.utils": ("HpmStruct", "torch_version_is_greater_or_equal"),
    "super_gradients.training.utils.bbox_utils": ("batch_distance2bbox",),
    "super_gradients.training.utils.utils": ("infer_model_dtype", "infer_model_device", "HpmStruct"),
    "super_gradients.module_interfaces": ("SupportsReplaceNumClasses", "AbstractPoseEstimationPostPredictionCallback", "SupportsInputShapeCheck"),
    "super_gradients.module_interfaces.supports_input_shape_check": ("SupportsInputShapeCheck",),
    "super_gradients.training.models": ("CustomizableDetector",),
}


class _PlaceholderMeta(type):
    def __getattr__(cls, name):   # class-level access, e.g. an enum member used as a default argument
        if name.startswith("__"):
            raise AttributeError(name)
        return _Placeholder()


class _Placeholder(metaclass=_PlaceholderMeta):
    """Inert stand-in for anything else: callable, subclassable, usable as a decorator, attribute-transparent."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        if len(a) == 1 and not k and (isinstance(a[0], type) or callable(a[0])):
            return a[0]   # used as a decorator
        return _Placeholder()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Placeholder()


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        obj = _PlaceholderMeta(name, (_Placeholder,), {})
        setattr(self, name, obj)
        return obj


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    PREFIXES = ("super_gradients.",)

    def find_spec(self, fullname, path=None, target=None):
        if fullname.startswith(self.PREFIXES):
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        for name in _REAL.get(module.__name__, ()):
            setattr(module, name, getattr(_impl, name))


if not any(isinstance(f, _Finder) for f in sys.meta_path):
    sys.meta_path.append(_Finder())
