"""crowddynamics_b200 -- B200-native per-timestep agent update of crowddynamics (hot path only).

Host side is Python over a C-ABI shared library (``csrc/libcrowd_b200.so``, hand-written sm_100a CUDA).
There is no CPU fallback: importing the engine without the built library raises ``ExtensionMissing``.
"""
from .exceptions import CrowdDynamicsException, InvalidType, InvalidValue, DeviceError, ExtensionMissing  # noqa
from .structures import (agent_type_circular, agent_type_three_circle, obstacle_type_linear,  # noqa
                         is_model, AgentModelToType)

__version__ = '0.1.0'
