"""Minimal stand-in for rospy: parameters from a dict, publishers that record, no spinning.  Used by
tests/test_node_dropin_cpu.py to run the reference's scripts/taichislam_node.py unmodified."""
PARAMS = {}
PUBLISHED = []


def get_param(name, default=None):
    return PARAMS.get(name, default)


class Publisher:
    def __init__(self, topic, msg_type, queue_size=10, **kw):
        self.topic = topic

    def publish(self, msg):
        PUBLISHED.append((self.topic, msg))


class Subscriber:
    def __init__(self, topic, msg_type, callback=None, queue_size=10, **kw):
        self.topic, self.callback = topic, callback


class Time:
    @staticmethod
    def now():
        return 0.0


class Rate:
    def __init__(self, hz):
        pass

    def sleep(self):
        pass


def init_node(name, **kw):
    pass


def is_shutdown():
    return True
