class _V:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class Odometry:
    def __init__(self, x=0.0, y=0.0, z=0.0):
        self.pose = _V(pose=_V(position=_V(x=x, y=y, z=z)))


class OccupancyGrid:
    def __init__(self):
        self.header = _V(stamp=None, frame_id=None)
        self.info = _V(resolution=None, width=None, height=None,
                       origin=_V(orientation=_V(x=0, y=0, z=0, w=1), position=_V(x=0, y=0, z=0)))
        self.data = None
