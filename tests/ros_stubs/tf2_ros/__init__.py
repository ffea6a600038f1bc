"""Stand-in for tf2_ros: a Buffer whose lookup_transform returns the pose registered for a stamp."""


class _V:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class Buffer:
    POSES = {}                               # stamp -> (translation xyz, quaternion xyzw)

    def lookup_transform(self, target, source, stamp, timeout=None):
        t, q = Buffer.POSES[stamp]
        return _V(transform=_V(translation=_V(x=t[0], y=t[1], z=t[2]), rotation=_V(x=q[0], y=q[1], z=q[2], w=q[3])))


class TransformListener:
    def __init__(self, buffer):
        self.buffer = buffer
