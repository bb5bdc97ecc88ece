"""Run the REFERENCE's own Python (utils.py and Env of
environment_stage_1_nobonus.py) in this container, without ROS / Gazebo.

Only usable where /root/reference exists (the build container).  Used by
tests/gen_golden.py to produce the committed fixtures under tests/golden/, and
by tests that are skipped when the reference is absent (e.g. on the GPU box).

How: the reference modules import rospy, tf, shapely and ROS message types.
Those are replaced by small stand-ins in sys.modules:
  * rospy / message types : inert objects; rospy.get_param serves the YAML values
  * tf.transformations.euler_from_quaternion : the standard yaw formula
  * shapely : a minimal restatement of exactly the calls the reference makes
      - Point.buffer(r).boundary : GEOS' default 64-gon (vertices at -k*pi/32)
      - ring.intersection(LineString) : exact segment/segment intersections
      - Polygon(...).intersection/union(...).area for axis-aligned boxes,
        Polygon.contains(Point) for convex quads
The reference source itself is executed UNMODIFIED except for two in-memory
text patches: the Python-2 integer division that Python 3 would turn into a
float list index (ENV:577 `len(x) / 2` -> `len(x) // 2`) and the hard-coded
observation offsets of compute_reward (ENV:1047-1048 `state[359]`, `state[360]`
-> `state[scan_ranges - 1]`, `state[scan_ranges]`; identical for 360 samples).
Nothing is copied into this repo.
Physics the reference gets from Gazebo (pose, twist, scan) is injected by the
caller each step.
"""
from __future__ import annotations

import importlib.util
import math
import os
import sys
import types

REF_SRC = "/root/reference/turtlebot3_rl_sim/src"


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_SRC, "utils.py"))


# --------------------------------------------------------------------------- stubs
class _Inert:
    def __init__(self, *a, **k):
        pass

    def __getattr__(self, n):
        return _Inert()

    def __call__(self, *a, **k):
        return _Inert()


class _XYZ:
    def __init__(self, x=0.0, y=0.0, z=0.0):
        self.x, self.y, self.z = x, y, z


class _Quat:
    def __init__(self, yaw=0.0):
        self.x, self.y, self.z, self.w = 0.0, 0.0, math.sin(yaw / 2.0), math.cos(yaw / 2.0)


class _Twist:
    def __init__(self):
        self.linear, self.angular = _XYZ(), _XYZ()


class _Pose:
    def __init__(self):
        self.position, self.orientation = _XYZ(), _Quat()
        self.x = self.y = self.z = 0.0     # Env.__init__ does `self.position = Pose()` then reads .x/.y


def _euler_from_quaternion(q):
    x, y, z, w = q
    roll = math.atan2(2.0 * (w * x + y * z), 1.0 - 2.0 * (x * x + y * y))
    pitch = math.asin(max(-1.0, min(1.0, 2.0 * (w * y - z * x))))
    yaw = math.atan2(2.0 * (w * z + x * y), 1.0 - 2.0 * (y * y + z * z))
    return roll, pitch, yaw


# -- minimal shapely ---------------------------------------------------------------
class _SPoint:
    def __init__(self, x, y):
        self.x, self.y = float(x), float(y)
        self.coords = [(self.x, self.y)]

    def buffer(self, r):
        # GEOS OffsetCurveBuilder: quadrantSegments=8?  shapely's default resolution=16 -> 64 segments,
        # fillet from angle 0 going CLOCKWISE
        pts = [(self.x + r * math.cos(-k * math.pi / 32.0), self.y + r * math.sin(-k * math.pi / 32.0))
               for k in range(64)]
        return _SPolygon(pts)

    def __str__(self):
        return "POINT (%r %r)" % (self.x, self.y)


class _SEmpty:
    def __str__(self):
        return "LINESTRING EMPTY"


class _SMultiPoint:
    def __init__(self, pts):
        self.geoms = [_SPoint(*p) for p in pts]

    def __str__(self):
        return "MULTIPOINT (...)"


class _SLineString:
    def __init__(self, coords):
        self.coords = [(float(a), float(b)) for a, b in coords]


def _seg_intersect(p, q, a, b):
    """Intersection point of segments pq and ab (None if disjoint / parallel)."""
    rx, ry = q[0] - p[0], q[1] - p[1]
    sx, sy = b[0] - a[0], b[1] - a[1]
    den = rx * sy - ry * sx
    if den == 0.0:
        return None
    t = ((a[0] - p[0]) * sy - (a[1] - p[1]) * sx) / den
    u = ((a[0] - p[0]) * ry - (a[1] - p[1]) * rx) / den
    if 0.0 <= t <= 1.0 and 0.0 <= u <= 1.0:
        return (p[0] + t * rx, p[1] + t * ry)
    return None


class _SRing:
    def __init__(self, pts):
        self.pts = list(pts)

    def intersection(self, line):
        p, q = line.coords[0], line.coords[1]
        hits = []
        n = len(self.pts)
        for k in range(n):
            h = _seg_intersect(p, q, self.pts[k], self.pts[(k + 1) % n])
            if h is not None and not any(abs(h[0] - o[0]) < 1e-12 and abs(h[1] - o[1]) < 1e-12 for o in hits):
                hits.append(h)
        if not hits:
            return _SEmpty()
        if len(hits) == 1:
            return _SPoint(*hits[0])          # a Point has no .geoms (shapely 1.x)
        hits.sort(key=lambda h: (h[0] - p[0]) ** 2 + (h[1] - p[1]) ** 2)
        return _SMultiPoint(hits)


class _SArea:
    def __init__(self, area):
        self.area = area


class _SPolygon:
    def __init__(self, pts):
        self.pts = [(float(a), float(b)) for a, b in pts]
        self.boundary = _SRing(self.pts)
        xs, ys = [p[0] for p in self.pts], [p[1] for p in self.pts]
        self._box = (min(xs), min(ys), max(xs), max(ys))
        a = 0.0
        for k in range(len(self.pts)):
            x0, y0 = self.pts[k]
            x1, y1 = self.pts[(k + 1) % len(self.pts)]
            a += x0 * y1 - x1 * y0
        self.area = abs(a) / 2.0

    # only ever called on axis-aligned squares (utils._get_bounding_box)
    def intersection(self, other):
        ax0, ay0, ax1, ay1 = self._box
        bx0, by0, bx1, by1 = other._box
        w, h = min(ax1, bx1) - max(ax0, bx0), min(ay1, by1) - max(ay0, by0)
        return _SArea(w * h if (w > 0 and h > 0) else 0.0)

    def union(self, other):
        return _SArea(self.area + other.area - self.intersection(other).area)

    def contains(self, pt):
        sign = 0
        n = len(self.pts)
        for k in range(n):
            x0, y0 = self.pts[k]
            x1, y1 = self.pts[(k + 1) % n]
            c = (x1 - x0) * (pt.y - y0) - (y1 - y0) * (pt.x - x0)
            if c == 0:
                return False
            s = 1 if c > 0 else -1
            if sign == 0:
                sign = s
            elif s != sign:
                return False
        return True


DEFAULT_PARAMS = {   # configs/turtlebot3_world.yaml:1-18
    "/turtlebot3/linear_forward_speed": 0.5, "/turtlebot3/linear_turn_speed": 0.05, "/turtlebot3/angular_speed": 0.3,
    "/turtlebot3/scan_ranges": 360, "/turtlebot3/max_scan_range": 0.6, "/turtlebot3/min_scan_range": 0.12,
    "/turtlebot3/desired_pose/x": -1.0, "/turtlebot3/desired_pose/y": 1.0, "/turtlebot3/desired_pose/z": 0.0,
    "/turtlebot3/starting_pose/x": 0.75, "/turtlebot3/starting_pose/y": -0.75, "/turtlebot3/starting_pose/z": 0.0,
}


class _Clock:
    """time.time()/time.sleep() stand-in: sleep advances the clock and lets the
    injected physics run (in the reference, Gazebo moves the robot during the
    0.15 s sleep of Env.step and the odom callback updates self.position)."""

    def __init__(self):
        self.now = 1000.0
        self.on_sleep = None

    def time(self):
        return self.now

    def sleep(self, dt):
        self.now += dt
        if self.on_sleep is not None:
            self.on_sleep(dt)


class Reference:
    """Loads the reference modules once with the stand-ins installed."""

    def __init__(self, params=None):
        if not reference_available():
            raise RuntimeError("/root/reference is not present")
        self.params = dict(DEFAULT_PARAMS)
        if params:
            self.params.update(params)
        self.clock = _Clock()
        self.scan_source = None                     # callable -> object with .ranges
        saved = {k: sys.modules.get(k) for k in list(sys.modules)}
        try:
            self._install()
            self.utils = self._load("utils", os.path.join(REF_SRC, "utils.py"), {})
            sys.modules["utils"] = self.utils
            self.envmod = self._load("environment_stage_1_nobonus",
                                     os.path.join(REF_SRC, "environment_stage_1_nobonus.py"),
                                     {"_center_item = len(segmented_scan_object_types_2d[i]) / 2":
                                      "_center_item = len(segmented_scan_object_types_2d[i]) // 2",
                                      # compute_reward hard-codes the 360-sample layout (ENV:1047-1048);
                                      # generalised so the 37-sample config can run (identical for 360)
                                      "current_heading = state[359]": "current_heading = state[self.scan_ranges - 1]",
                                      "current_distance = state[360]": "current_distance = state[self.scan_ranges]"})
            # the reference's first environment (environment_stage_1_original.py), unmodified
            self.envmod_original = self._load("environment_stage_1_original",
                                              os.path.join(REF_SRC, "environment_stage_1_original.py"), {})
            # original:284-286 appends every pose to a CSV under the ROS package path: not part of the path under test
            self.utils.record_data = lambda *a, **k: None
        finally:
            for k in ("rospy", "tf", "rospkg", "tf.transformations", "shapely", "shapely.geometry", "shapely.geometry.polygon",
                      "visualization_msgs", "visualization_msgs.msg", "geometry_msgs", "geometry_msgs.msg",
                      "sensor_msgs", "sensor_msgs.msg", "nav_msgs", "nav_msgs.msg", "std_srvs", "std_srvs.srv",
                      "utils"):
                if saved.get(k) is None:
                    sys.modules.pop(k, None)
                else:
                    sys.modules[k] = saved[k]

    def _install(self):
        def mod(name, **attrs):
            m = types.ModuleType(name)
            m.__dict__.update(attrs)
            sys.modules[name] = m
            return m
        ref = self

        def wait_for_message(topic, *a, **k):
            return ref.scan_source()

        mod("rospy", Publisher=_Inert, Subscriber=_Inert, ServiceProxy=_Inert, on_shutdown=lambda f: None,
            get_param=lambda k: ref.params[k], loginfo=lambda *a: None, logwarn=lambda *a: None,
            wait_for_service=lambda *a, **k: None, wait_for_message=wait_for_message,
            ServiceException=Exception, get_name=lambda: "ref", init_node=lambda *a, **k: None,
            Time=_Inert(), Duration=_Inert, Rate=_Inert, is_shutdown=lambda: False)
        mod("tf")
        mod("tf.transformations", euler_from_quaternion=_euler_from_quaternion)
        mod("rospkg", RosPack=lambda: types.SimpleNamespace(get_path=lambda name: "/tmp"))
        mod("shapely")
        mod("shapely.geometry", LineString=_SLineString, Point=_SPoint)
        mod("shapely.geometry.polygon", Polygon=_SPolygon)
        mod("visualization_msgs")
        mod("visualization_msgs.msg", Marker=_Inert)
        mod("geometry_msgs")
        mod("geometry_msgs.msg", Twist=_Twist, Pose=_Pose, Point=_XYZ, PointStamped=_Inert)
        mod("sensor_msgs")
        mod("sensor_msgs.msg", LaserScan=_Inert)
        mod("nav_msgs")
        mod("nav_msgs.msg", Odometry=_Inert)
        mod("std_srvs")
        mod("std_srvs.srv", Empty=_Inert)

    def _load(self, name, path, patches):
        import warnings
        with open(path) as f:
            src = f.read()
        for old, new in patches.items():
            assert src.count(old) == 1, "patch target not found exactly once: %r" % old
            src = src.replace(old, new)
        m = types.ModuleType(name)
        m.__file__ = path
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", SyntaxWarning)
            code = compile(src, path, "exec")
        exec(code, m.__dict__)
        # route wall-clock calls through the injected clock; silence the reference's prints
        fake_time = types.SimpleNamespace(time=self.clock.time, sleep=self.clock.sleep)
        if "time" in m.__dict__:
            m.__dict__["time"] = fake_time
        m.__dict__["print"] = lambda *a, **k: None
        return m

    # ---- Env with injected physics --------------------------------------------------
    def make_env(self, action_dim=2, max_step=1000, k_obstacle_count=8):
        env = self.envmod.Env(action_dim=action_dim, max_step=max_step)
        env.k_obstacle_count = k_obstacle_count          # hard-coded 8 at ENV:55
        # Env.__init__ leaves position = Pose(); the odom callback would overwrite these four
        env.position = _XYZ()
        env.orientation = _Quat(0.0)
        env.linear_twist = _XYZ()
        env.angular_twist = _XYZ()
        return env

    def make_env_original(self, action_dim=2, max_step=1000):
        env = self.envmod_original.Env(action_dim=action_dim, max_step=max_step)
        env.position = _XYZ()
        env.orientation = _Quat(0.0)
        return env

    @staticmethod
    def set_odom(env, x, y, yaw, v, w):
        """What Env.get_odometry (ENV:239-243) stores from an Odometry message."""
        env.position = _XYZ(x, y, 0.0)
        env.orientation = _Quat(yaw)
        env.linear_twist = _XYZ(v, 0.0, 0.0)
        env.angular_twist = _XYZ(0.0, 0.0, w)


class Scan:
    def __init__(self, ranges):
        self.ranges = list(ranges)
