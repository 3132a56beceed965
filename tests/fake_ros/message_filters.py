class Subscriber:
    def __init__(self, topic, msg_type, queue_size=10, **kw):
        self.topic = topic


class ApproximateTimeSynchronizer:
    def __init__(self, subs, queue_size, slop=0.1):
        self.subs, self.callback = subs, None

    def registerCallback(self, cb):
        self.callback = cb
