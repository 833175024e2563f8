from .exporter_utils import sample_volume, write_ply  # noqa: F401
