class DroneTraj:
    def __init__(self, drone_id=1, frame_ids=(), poses=()):
        self.drone_id, self.frame_ids, self.poses = drone_id, list(frame_ids), list(poses)


class _Odom:
    def __init__(self, pose):
        self.pose = type("P", (), {"pose": pose})()


class VIOFrame:
    def __init__(self, frame_id, is_keyframe, pose, extrinsics):
        self.frame_id, self.is_keyframe = frame_id, is_keyframe
        self.odom = _Odom(pose)
        self.extrinsics = list(extrinsics)
