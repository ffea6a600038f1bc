class _V:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class PointCloud2:
    def __init__(self, points=None, frame_id="os_sensor", stamp=0):
        self.points = points
        self.header = _V(frame_id=frame_id, stamp=stamp)
