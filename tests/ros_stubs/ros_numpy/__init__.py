from . import point_cloud2  # noqa: F401
