"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference hot-path file.

This module exists to (a) validate ``oracle/gd_oracle.py`` against the real
reference and (b) generate the golden fixtures under ``tests/golden/``.  It only
reads ``/root/reference`` inside the build container; on the GPU box, which has no such
path, it finds the byte-identical copy of the one file that ``oracle/build_ref.py`` stages
under ``oracle/_ref/`` (git-ignored).  Only ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs use it there (the reference's own CPU path, timed beside ours).

The reference file ``mmdet3d_gaussian/models/losses/gaussian_distance_loss.py``
imports exactly two upstream symbols (lines 3-4):

* ``mmdet.models.builder.LOSSES``            -- a registry with a
  ``register_module()`` decorator (``gaussian_distance_loss.py:251``)
* ``mmdet.models.losses.utils.weighted_loss`` -- decorator applied at
  ``gaussian_distance_loss.py:42,109,144,189,201,214,227``

mmdet is NOT installed in this image (and is un-pinned by the reference:
``setup.py`` has no ``install_requires``), so both are supplied by the stub
below.  ``weighted_loss`` restates mmdet 2.x ``models/losses/utils.py``
(``weight_reduce_loss``): ``loss*=weight``; ``avg_factor is None`` ->
none/mean/sum; ``avg_factor`` given -> ``mean``: ``sum()/avg_factor``,
``none``: unchanged, ``sum``: ValueError.  (Some mmdet releases add
``finfo(float32).eps`` to ``avg_factor``; <=1.2e-7 relative, inside tolerance,
and not reproduced here.)
"""
import functools
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('GD_REFERENCE_ROOT', '/root/reference')
REFERENCE_FILE = os.path.join(
    REFERENCE_ROOT, 'mmdet3d_gaussian', 'models', 'losses',
    'gaussian_distance_loss.py')


# On the GPU box /root/reference does not exist; oracle/build_ref.py stages a byte-identical
# copy of the ONE file under oracle/_ref/ (git-ignored) and this loader falls back to it.
STAGED_FILE = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref',
                           'gaussian_distance_loss.py')


def reference_file():
    """Path of the reference loss file: the live checkout, else the staged copy, else None."""
    if os.path.isfile(REFERENCE_FILE):
        return REFERENCE_FILE
    if os.path.isfile(STAGED_FILE):
        from . import build_ref
        if build_ref.staged_is_intact():
            return STAGED_FILE
    return None


def reference_available():
    return reference_file() is not None


class _StubRegistry:
    """Minimal stand-in for mmcv's Registry: name -> class, decorator API."""

    def __init__(self, name):
        self.name = name
        self.module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def _register(cls):
            self.module_dict[name or cls.__name__] = cls
            return cls
        if module is not None:
            return _register(module)
        return _register

    def get(self, key):
        return self.module_dict.get(key)

    def build(self, cfg):
        cfg = dict(cfg)
        return self.module_dict[cfg.pop('type')](**cfg)


def _weight_reduce_loss(loss, weight=None, reduction='mean', avg_factor=None):
    """mmdet 2.x ``weight_reduce_loss`` semantics."""
    if weight is not None:
        loss = loss * weight
    if avg_factor is None:
        if reduction == 'mean':
            return loss.mean()
        if reduction == 'sum':
            return loss.sum()
        return loss
    if reduction == 'mean':
        return loss.sum() / avg_factor
    if reduction != 'none':
        raise ValueError('avg_factor can not be used with reduction="sum"')
    return loss


def _weighted_loss(loss_func):
    @functools.wraps(loss_func)
    def wrapper(pred, target, weight=None, reduction='mean', avg_factor=None,
                **kwargs):
        loss = loss_func(pred, target, **kwargs)
        return _weight_reduce_loss(loss, weight, reduction, avg_factor)
    return wrapper


def _install_stub_mmdet():
    if 'mmdet.models.builder' in sys.modules and hasattr(
            sys.modules['mmdet.models.builder'], '_gd_stub'):
        return
    names = ['mmdet', 'mmdet.models', 'mmdet.models.builder',
             'mmdet.models.losses', 'mmdet.models.losses.utils']
    mods = {n: types.ModuleType(n) for n in names}
    for n in names:
        if '.' in n:
            parent, child = n.rsplit('.', 1)
            setattr(mods[parent], child, mods[n])
    mods['mmdet.models.builder'].LOSSES = _StubRegistry('loss')
    mods['mmdet.models.builder']._gd_stub = True
    mods['mmdet.models.losses.utils'].weighted_loss = _weighted_loss
    mods['mmdet.models.losses.utils'].weight_reduce_loss = _weight_reduce_loss
    sys.modules.update(mods)


_CACHED = None


def load_reference():
    """Import the reference loss file by path, unmodified; returns the module."""
    global _CACHED
    if _CACHED is not None:
        return _CACHED
    path = reference_file()
    if path is None:
        raise FileNotFoundError(
            f'{REFERENCE_FILE} not found and no intact staged copy under oracle/_ref/ '
            f'(python oracle/build_ref.py in the build container)')
    _install_stub_mmdet()
    spec = importlib.util.spec_from_file_location(
        '_gd_reference_loss', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _CACHED = mod
    return mod


# ---------------------------------------------------------------------------
# CenterPointBBoxYawCoder: the decoder the reference runs
# in front of the loss in CenterGDHead.loss (gd_centerpoint_head.py:421-423).
# ---------------------------------------------------------------------------
CODER_DIR = os.path.join(REFERENCE_ROOT, 'mmdet3d_gaussian', 'core', 'bbox', 'coders')
_CODER = None


def load_reference_center_coder():
    """Import ``core/bbox/coders/centerpoint_bbox_coders.py`` and
    ``centerpoint_bbox_yaw_coders.py`` by path, unmodified, under a stub
    ``mmdet.core.bbox`` (``BaseBBoxCoder`` = plain base class, ``BBOX_CODERS`` =
    registry with ``register_module()``); returns ``CenterPointBBoxYawCoder``."""
    global _CODER
    if _CODER is not None:
        return _CODER
    if not os.path.isdir(CODER_DIR):
        raise FileNotFoundError(f'{CODER_DIR} not found: build container only')
    _install_stub_mmdet()
    for name in ('mmdet.core', 'mmdet.core.bbox', 'mmdet.core.bbox.builder'):
        if name not in sys.modules:
            mod = types.ModuleType(name)
            sys.modules[name] = mod
            parent, child = name.rsplit('.', 1)
            setattr(sys.modules[parent], child, mod)
    sys.modules['mmdet.core.bbox'].BaseBBoxCoder = type('BaseBBoxCoder', (), {})
    sys.modules['mmdet.core.bbox.builder'].BBOX_CODERS = _StubRegistry('bbox_coder')
    pkg = types.ModuleType('_gd_ref_coders')
    pkg.__path__ = [CODER_DIR]               # a package whose __init__ is NOT executed
    sys.modules['_gd_ref_coders'] = pkg

    def _load(stem):
        spec = importlib.util.spec_from_file_location(
            f'_gd_ref_coders.{stem}', os.path.join(CODER_DIR, stem + '.py'))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = mod
        spec.loader.exec_module(mod)
        return mod
    pkg.CenterPointBBoxCoderRev = _load('centerpoint_bbox_coders').CenterPointBBoxCoderRev
    _CODER = _load('centerpoint_bbox_yaw_coders').CenterPointBBoxYawCoder
    return _CODER


# ---------------------------------------------------------------------------
# SimOTABEVAssigner: only its pure-torch method
# dynamic_k_matching (sim_ota_3d_assigner.py:184-211) is on our path; the module's imports
# (mmdet assigner base classes, mmdet3d ops / box structures) are stubbed.
# ---------------------------------------------------------------------------
SIMOTA_FILE = os.path.join(REFERENCE_ROOT, 'mmdet3d_gaussian', 'core', 'bbox', 'assigners',
                           'sim_ota_3d_assigner.py')
_SIMOTA = None


def load_reference_simota():
    """Import ``core/bbox/assigners/sim_ota_3d_assigner.py`` by path, unmodified; returns the
    class ``SimOTABEVAssigner``."""
    global _SIMOTA
    if _SIMOTA is not None:
        return _SIMOTA
    if not os.path.isfile(SIMOTA_FILE):
        raise FileNotFoundError(f'{SIMOTA_FILE} not found: build container only')
    _install_stub_mmdet()
    stubs = ['mmdet.core', 'mmdet.core.bbox', 'mmdet.core.bbox.assigners', 'mmdet.core.bbox.builder',
             'mmdet3d', 'mmdet3d.ops', 'mmdet3d.core', 'mmdet3d.core.bbox',
             'mmdet3d.core.bbox.structures', 'mmdet3d.core.bbox.structures.lidar_box3d']
    for name in stubs:
        if name not in sys.modules:
            mod = types.ModuleType(name)
            sys.modules[name] = mod
            if '.' in name:
                parent, child = name.rsplit('.', 1)
                setattr(sys.modules[parent], child, mod)
    sys.modules['mmdet.core.bbox.assigners'].BaseAssigner = type('BaseAssigner', (), {})
    sys.modules['mmdet.core.bbox.assigners'].AssignResult = type('AssignResult', (), {})
    if not hasattr(sys.modules['mmdet.core.bbox.builder'], 'BBOX_ASSIGNERS'):
        sys.modules['mmdet.core.bbox.builder'].BBOX_ASSIGNERS = _StubRegistry('bbox_assigner')
    sys.modules['mmdet3d.ops'].points_in_boxes_all = None
    sys.modules['mmdet3d.core.bbox.structures.lidar_box3d'].LiDARInstance3DBoxes = None
    spec = importlib.util.spec_from_file_location('_gd_reference_simota', SIMOTA_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _SIMOTA = mod.SimOTABEVAssigner
    return _SIMOTA
