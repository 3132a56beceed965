class Header:
    def __init__(self, frame_id="", stamp=None, seq=0):
        self.frame_id, self.stamp, self.seq = frame_id, stamp, seq
