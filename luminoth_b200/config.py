"""YAML config surface of the predict path.

Mirrors ``luminoth/utils/config.py``: ``get_config`` :14-22,
``load_config_files`` :25-45, ``merge_into`` :113-148 (type-compat check
:73-92, ``_replace`` meta-key :95-110), ``parse_override`` :151-171,
``parse_config_value`` :174-196, ``get_model_config`` :213-225,
``override_config_params`` :228-232.  Same names, argument meaning and error
behaviour (``ValueError`` on incompatible types / malformed overrides).
``easydict`` is not installed here, so ``Config`` is a minimal attribute-dict
with the same observable behaviour for this path (attribute + item access,
recursive wrapping of nested dicts).
"""
import os

import yaml

REPLACE_KEY = '_replace'
_CONFIG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'configs')
MODEL_TYPES = ('fasterrcnn', 'ssd')     # luminoth/models/models.py:7-10


class Config(dict):
    """dict with attribute access; nested dicts (also inside lists) are wrapped."""

    def __init__(self, d=None, **kwargs):
        super().__init__()
        d = dict(d or {})
        d.update(kwargs)
        for k, v in d.items():
            self[k] = v

    @classmethod
    def _wrap(cls, v):
        if isinstance(v, dict) and not isinstance(v, Config):
            return cls(v)
        if isinstance(v, (list, tuple)):
            return type(v)(cls._wrap(x) for x in v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, self._wrap(v))

    def __setattr__(self, k, v):
        self[k] = v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def update(self, *a, **kw):
        for k, v in dict(*a, **kw).items():
            self[k] = v

    def copy(self):
        return Config(self)

    def to_dict(self):
        def conv(v):
            if isinstance(v, dict):
                return {k: conv(x) for k, x in v.items()}
            if isinstance(v, (list, tuple)):
                return [conv(x) for x in v]
            return v
        return conv(self)


def get_model_type(model_type):
    """``luminoth/models/models.py:13-17`` (ValueError on unknown types)."""
    if model_type not in MODEL_TYPES:
        raise ValueError('"{}" is not a valid model_type'.format(model_type))
    return model_type


def load_config_files(filename_or_filenames, warn_overwrite=True):
    if isinstance(filename_or_filenames, (list, tuple)):
        filenames = filename_or_filenames
    else:
        filenames = [filename_or_filenames]
    config = Config({})
    for filename in filenames:
        with open(filename) as f:
            new_config = Config(yaml.safe_load(f) or {})
        config = merge_into(new_config, config, overwrite=True)
    return config


def get_base_config(model_type):
    return load_config_files([os.path.join(_CONFIG_DIR, get_model_type(model_type) + '.yml')])


def types_compatible(new_value, base_value):
    if base_value is None:
        return True
    if new_value is None or new_value is False:
        return True
    if isinstance(new_value, str) and isinstance(base_value, str):
        return True
    return isinstance(new_value, type(base_value))


def should_replace(new_config, base_config, key):
    try:
        base_replace = base_config[key][REPLACE_KEY]
    except (KeyError, TypeError):
        base_replace = None
    try:
        new_replace = new_config[key][REPLACE_KEY]
    except (KeyError, TypeError):
        new_replace = None
    if new_replace:
        return True
    return bool(new_replace is None and base_replace)


def merge_into(new_config, base_config, overwrite=False, warn_overwrite=False):
    if not isinstance(new_config, Config):
        return
    for key, value in new_config.items():
        if not types_compatible(value, base_config.get(key)):
            raise ValueError('Incorrect type "{}" for key "{}". Must be "{}"'.format(
                type(value), key, type(base_config.get(key))))
        if isinstance(value, dict):
            if should_replace(new_config, base_config, key):
                base_config[key] = value
            else:
                base_config[key] = merge_into(
                    new_config[key], base_config.get(key) or Config({}),
                    overwrite=overwrite, warn_overwrite=warn_overwrite)
        else:
            if base_config.get(key) is None:
                base_config[key] = value
            elif overwrite:
                base_config[key] = value
    return base_config


def parse_config_value(value):
    low = value.lower()
    if low == 'none':
        return None
    if low == 'true':
        return True
    if low == 'false':
        return False
    try:
        return int(value)
    except ValueError:
        pass
    try:
        return float(value)
    except ValueError:
        pass
    return value


def parse_override(override_options):
    if not override_options:
        return {}
    override_dict = {}
    for option in override_options:
        key_value = option.split('=')
        if len(key_value) != 2:
            raise ValueError('Invalid override option "{}"'.format(option))
        key, value = key_value
        nested = key.split('.')
        cur = override_dict
        for k in nested[:-1]:
            cur = cur.setdefault(k, {})
        cur[nested[-1]] = parse_config_value(value)
    return override_dict


def cleanup_config(config):
    config.pop(REPLACE_KEY, None)
    for k in config:
        if isinstance(config[k], dict):
            cleanup_config(config[k])
    return config


def get_model_config(base_config, custom_config, override_params):
    config = Config(base_config.copy())
    if custom_config:
        config = merge_into(Config(custom_config), config, overwrite=True)
    if override_params:
        config = merge_into(Config(parse_override(override_params)), config, overwrite=True)
    return cleanup_config(config)


def override_config_params(config, params):
    return merge_into(Config(parse_override(params)), config, overwrite=True)


def get_config(config_files, override_params=None):
    custom_config = load_config_files(config_files)
    base = get_base_config(custom_config['model']['type'])
    return get_model_config(base, custom_config, override_params)


def default_config(model_type, override_params=None):
    """Convenience (not in the reference): base config of a model type + overrides."""
    return get_model_config(get_base_config(model_type), None, override_params)


def set_prediction_filters(config, min_prob=None, max_detections=None):
    """The config mutations every caller of ``PredictorNetwork`` applies BEFORE building it:

    * ``lumi predict`` (``predict.py:246-259``): ``max_detections`` (CLI default 100) overwrites
      ``rcnn.proposals.total_max_detections`` -- or ``rpn.proposals.post_nms_top_n`` when ``with_rcnn`` is off --
      and ``min_prob`` (CLI default 0.5) overwrites ``min_prob_threshold``;
    * ``Detector`` (``tasks.py:64-67``) forces ``min_prob_threshold = 0.0`` and filters in Python;
    * the web server (``tools/server/web.py:97-100``) uses 0.01.

    ``None`` leaves the respective value untouched.  Unknown model types raise ``ValueError`` like the reference."""
    mtype = config.model.type
    if mtype == 'fasterrcnn':
        if max_detections is not None:
            if config.model.network.get('with_rcnn', False):
                config.model.rcnn.proposals.total_max_detections = max_detections
            else:
                config.model.rpn.proposals.post_nms_top_n = max_detections
        if min_prob is not None:
            config.model.rcnn.proposals.min_prob_threshold = min_prob
    elif mtype == 'ssd':
        if max_detections is not None:
            config.model.proposals.total_max_detections = max_detections
        if min_prob is not None:
            config.model.proposals.min_prob_threshold = min_prob
    else:
        raise ValueError("Model type '{}' not supported".format(mtype))
    return config

