from .fruit_datamanager import FruitDataManager, FruitDataManagerConfig, get_corners_of_aabb, sample_surface_points  # noqa: F401
