"""Mirror of the reference's ``src.data`` for the hot path (src/data/__init__.py:7-15): AddSensorNoise is
re-implemented, plus the batched GPU form of the dataset's per-sample preprocessing (ingest_tiles);
file I/O and the DataLoader stay with the reference."""
from .transforms import AddSensorNoise
from .ingest import export_tiles, ingest_tiles

__all__ = ["AddSensorNoise", "ingest_tiles", "export_tiles"]
