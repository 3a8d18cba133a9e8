"""In-memory stand-in for the `redis` module (TEST INFRASTRUCTURE ONLY).

The reference learner talks to a localhost Redis server purely as IPC
(/root/reference/LunarLander_Distributed_SAC/src/learner.py:32-36,
 .../replay_buffer.py:21,45-51, .../logger.py:12,35-41).  `redis` is not
installed in this image, and `redis.StrictRedis(host='localhost')` is evaluated
as a default argument at import time, so the stub has to be in `sys.modules`
before any reference module is imported.  Only the calls the learner-side code
makes are implemented: scan_iter/delete/set/get/rpush/llen and
pipeline().lrange/ltrim/execute.
"""
import sys
import threading
import types


class _Pipeline:
    def __init__(self, server):
        self._server = server
        self._ops = []

    def lrange(self, key, start, stop):
        self._ops.append(("lrange", key, start, stop))
        return self

    def ltrim(self, key, start, stop):
        self._ops.append(("ltrim", key, start, stop))
        return self

    def execute(self):
        out = []
        with self._server._lock:
            for op, key, start, stop in self._ops:
                lst = self._server._lists.get(key, [])
                if op == "lrange":
                    if stop == -1:
                        out.append(list(lst[start:]))
                    else:
                        out.append(list(lst[start:stop + 1]))
                else:  # ltrim: the reference uses (-1, 0) == "drop everything"
                    n = len(lst)
                    s = start + n if start < 0 else start
                    e = stop + n if stop < 0 else stop
                    self._server._lists[key] = lst[s:e + 1] if s <= e else []
                    out.append(True)
        self._ops = []
        return out


class StrictRedis:
    _shared = {}

    def __init__(self, host="localhost", **_kw):
        # one shared keyspace per host, like a real server
        st = StrictRedis._shared.setdefault(host, {"kv": {}, "lists": {}, "lock": threading.Lock()})
        self._kv, self._lists, self._lock = st["kv"], st["lists"], st["lock"]

    def scan_iter(self):
        with self._lock:
            return list(self._kv.keys()) + list(self._lists.keys())

    def delete(self, key):
        with self._lock:
            self._kv.pop(key, None)
            self._lists.pop(key, None)

    def set(self, key, value):
        with self._lock:
            self._kv[key] = value

    def get(self, key):
        with self._lock:
            return self._kv.get(key)

    def rpush(self, key, value):
        with self._lock:
            self._lists.setdefault(key, []).append(value)

    def llen(self, key):
        with self._lock:
            return len(self._lists.get(key, []))

    def pipeline(self):
        return _Pipeline(self)


Redis = StrictRedis


def install():
    """Put the stub into sys.modules as `redis` (idempotent)."""
    if "redis" not in sys.modules:
        mod = types.ModuleType("redis")
        mod.StrictRedis = StrictRedis
        mod.Redis = Redis
        sys.modules["redis"] = mod
    return sys.modules["redis"]
