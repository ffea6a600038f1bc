from . import msg, point_cloud2  # noqa: F401
