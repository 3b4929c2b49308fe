"""B200-native batched NB-IoT cell simulator behind the Gymnasium-style interface of
jjalcaraz-upct/nb-iot-environment. The compute path is CUDA only."""
from . import parameters  # noqa: F401
from .record import RECORD_DTYPE, STATE_NAMES, make_conf, record_to_info  # noqa: F401
from .controller import ControllerTables  # noqa: F401

__all__ = ['BatchedNBIoTEnv', 'PhiloxShim', 'parameters', 'RECORD_DTYPE', 'STATE_NAMES', 'make_conf', 'record_to_info', 'ControllerTables']


def __getattr__(name):
    # torch / the CUDA library are only needed by the environment itself
    if name == 'BatchedNBIoTEnv':
        from .env import BatchedNBIoTEnv
        return BatchedNBIoTEnv
    if name == 'PhiloxShim':
        from .philox_shim import PhiloxShim
        return PhiloxShim
    raise AttributeError(name)
