"""Mirror of the reference's ``src.data`` for the hot path (src/data/__init__.py:7-15): only AddSensorNoise is
re-implemented; dataset / dataloader I/O stays with the reference."""
from .transforms import AddSensorNoise

__all__ = ["AddSensorNoise"]
