"""luminoth_b200: B200-native (sm_100a) inference engine for the Faster R-CNN /
SSD predict path of tryolabs/luminoth, behind Luminoth's own
PredictorNetwork / config-YAML surface.  No CPU fallback."""
from .config import get_config, default_config, override_config_params  # noqa: F401

__version__ = '0.1'


def get_predictor(config, **kwargs):
    from .predicting import PredictorNetwork
    return PredictorNetwork(config, **kwargs)
