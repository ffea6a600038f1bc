"""Stand-in for rospy: parameters from a dict, publishers that record what is published."""
import time as _time

PARAMS = {}
PUBLISHERS = {}
LOG = []


def get_param(name, default=None):
    return PARAMS.get(name, default)


class Publisher:
    def __init__(self, topic, msg_type, queue_size=1):
        self.topic, self.messages = topic, []
        PUBLISHERS[topic] = self

    def publish(self, msg):
        import copy
        self.messages.append(copy.copy(msg))


class Subscriber:
    def __init__(self, topic, msg_type, callback, queue_size=1):
        self.topic, self.callback = topic, callback


class Duration:
    def __init__(self, secs=0.0):
        self.secs = secs


class Timer:
    def __init__(self, period, callback):
        self.period, self.callback = period, callback


class Time:
    @staticmethod
    def now():
        return _time.time()


def loginfo(msg):
    LOG.append(msg)


def init_node(name):
    pass


def is_shutdown():
    return True


def spin():
    pass


def on_shutdown(fn):
    pass
