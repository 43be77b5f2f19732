"""Registry hookup: the drop-in ``GDLoss`` must be reachable by the reference's
config key ``type='GDLoss'`` through mmdet's ``LOSSES`` registry
(reference ``gaussian_distance_loss.py:3,251``; built by ``build_loss`` at
``gd_anchor3d_head.py:60`` and ``gd_centerpoint_head.py:370``).

mmdet/mmcv are not installed in this image, so a local registry with the same
``register_module`` / ``build`` surface is always provided; when mmdet is
importable the class is additionally registered there with ``force=True`` so it
replaces the reference's eager implementation without a config change.
"""


class Registry:
    """The slice of mmcv's ``Registry`` that loss configs use."""

    def __init__(self, name):
        self.name = name
        self._modules = {}

    def register_module(self, name=None, force=False, module=None):
        def _register(cls):
            key = name or cls.__name__
            if key in self._modules and not force:
                raise KeyError(f'{key} is already registered in {self.name}')
            self._modules[key] = cls
            return cls
        if module is not None:
            return _register(module)
        return _register

    def get(self, key):
        return self._modules.get(key)

    def __contains__(self, key):
        return key in self._modules

    def build(self, cfg, default_args=None):
        if not isinstance(cfg, dict) or 'type' not in cfg:
            raise TypeError('cfg must be a dict containing the key "type"')
        args = dict(cfg)
        if default_args:
            for k, v in default_args.items():
                args.setdefault(k, v)
        obj_type = args.pop('type')
        cls = self._modules.get(obj_type) if isinstance(obj_type, str) else obj_type
        if cls is None:
            raise KeyError(f'{obj_type} is not in the {self.name} registry')
        return cls(**args)


LOSSES = Registry('loss')


def build_loss(cfg):
    """Same role as ``mmdet.models.builder.build_loss``."""
    return LOSSES.build(cfg)


def register_everywhere(cls):
    """Register ``cls`` locally and, if present, in mmdet's ``LOSSES``."""
    LOSSES.register_module(force=True)(cls)
    try:
        from mmdet.models.builder import LOSSES as MMDET_LOSSES
    except Exception:      # mmdet absent (this image) or broken: local registry only
        return cls
    try:
        MMDET_LOSSES.register_module(force=True)(cls)
    except TypeError:      # very old mmcv without ``force``
        MMDET_LOSSES._module_dict[cls.__name__] = cls
    return cls
