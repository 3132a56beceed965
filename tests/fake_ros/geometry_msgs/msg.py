class _V:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class Point(_V):
    pass


class Point32(_V):
    pass


class Quaternion(_V):
    pass


class Pose:
    def __init__(self, position=None, orientation=None):
        self.position = position or Point(x=0.0, y=0.0, z=0.0)
        self.orientation = orientation or Quaternion(x=0.0, y=0.0, z=0.0, w=1.0)


class PoseStamped:
    def __init__(self):
        self.pose = Pose()
