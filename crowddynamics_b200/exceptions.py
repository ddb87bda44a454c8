"""Exception types of the drop-in boundary.

Mirror crowddynamics/exceptions.py:10-22 of the reference.  ``MultiAgentProcess`` only catches
``CrowdDynamicsException`` (simulation/multiagent.py:89-94), so when the real package is importable our
exceptions *are* the reference's classes; otherwise equivalent stand-ins are defined.
"""
try:  # pragma: no cover - the reference package is not installable in the build image
    from crowddynamics.exceptions import CrowdDynamicsException, InvalidType, InvalidValue
except Exception:  # noqa
    class CrowdDynamicsException(Exception):
        """CrowdDynamics base exception."""

    class InvalidType(CrowdDynamicsException, TypeError):
        """Arguments to a CrowdDynamics function were of invalid type (e.g. unknown agent dtype)."""

    class InvalidValue(CrowdDynamicsException, ValueError):
        """Arguments to a CrowdDynamics function had an incorrect value."""


class DeviceError(CrowdDynamicsException, RuntimeError):
    """CUDA runtime / kernel failure reported by the C-ABI library."""


class ExtensionMissing(CrowdDynamicsException, ImportError):
    """libcrowd_b200.so is not built / cannot be loaded.  There is no CPU fallback."""
