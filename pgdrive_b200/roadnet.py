"""Analytic lanes and the road graph of a procedurally generated map (host side, float64).

Semantics follow the reference (cited per function); the structure does not: lanes are plain
slotted records, roads are ``(from_node, to_node)`` string pairs, and the graph is a thin wrapper
around an insertion-ordered dict whose iteration order defines the flat lane numbering that the
device tables use.

Reference: /root/reference/pgdrive/component/lane/{straight,circular,abs}_lane.py,
component/road/{road,road_network}.py, utils/scene_utils.py:40-136.
"""
import math

import numpy as np

# line types (constants.py LineType) and colours
NONE, BROKEN, CONTINUOUS, SIDE = "none", "broken", "continuous", "side"
GREY, YELLOW = "G", "Y"

DECO = ("decoration", "decoration_")  # constants.py:83-88

SIDEWALK_WIDTH = 3.0  # constants.py DrivableAreaProperty
SIDEWALK_LINE_DIST = 0.6


def norm2(x, y):
    """cutils_norm (cutils.pyx:147): spelt with ``**`` because C ``pow(x, 2)`` is not always the
    correctly rounded ``x * x`` and lane lengths feed ``int(length / 10)`` slot counts."""
    return math.sqrt(x**2 + y**2)


def wrap_to_pi(x):
    return ((x + np.pi) % (2 * np.pi)) - np.pi  # math_utils.py:32-33


class Lane:
    """Straight segment (kind 'S') or circular arc (kind 'C')."""
    __slots__ = (
        "kind", "sx", "sy", "ex", "ey", "dx", "dy", "length", "heading", "cx", "cy", "radius", "ph0", "ph1", "dir",
        "width", "line_types", "line_color", "speed_limit", "index"
    )
    DEFAULT_WIDTH = 4  # abs_lane.py:15

    def clone(self):
        c = Lane.__new__(Lane)
        for k in Lane.__slots__:
            v = getattr(self, k, None)
            setattr(c, k, list(v) if isinstance(v, list) else v)
        return c

    # -- constructors ------------------------------------------------------------------------------
    @staticmethod
    def straight(start, end, width, line_types=(BROKEN, BROKEN), speed_limit=1000):
        ln = Lane.__new__(Lane)
        ln.kind = "S"
        ln.sx, ln.sy = float(start[0]), float(start[1])
        ln.ex, ln.ey = float(end[0]), float(end[1])
        ln.width = width
        ln.line_types = line_types or [BROKEN, BROKEN]
        ln.line_color = [GREY, GREY]
        ln.speed_limit = speed_limit
        ln.index = None
        ln.cx = ln.cy = ln.radius = ln.ph0 = ln.ph1 = 0.0
        ln.dir = 0
        ln.refresh()
        ln.speed_limit = speed_limit
        return ln

    @staticmethod
    def arc(center, radius, ph0, ph1, clockwise, width, line_types=(BROKEN, BROKEN), speed_limit=1000):
        ln = Lane.__new__(Lane)
        ln.kind = "C"
        ln.cx, ln.cy = float(center[0]), float(center[1])
        ln.radius = radius
        ln.ph0, ln.ph1 = ph0, ph1
        ln.dir = 1 if clockwise else -1
        ln.width = width
        ln.line_types = line_types
        ln.line_color = [GREY, GREY]
        ln.speed_limit = speed_limit
        ln.index = None
        ln.dx = ln.dy = ln.heading = 0.0
        ln.refresh()
        return ln

    def refresh(self):
        """straight_lane.py:46-51 / circular_lane.py:41-44 (update_properties)."""
        if self.kind == "S":
            # StraightLane.update_properties re-runs AbstractLane.__init__, which resets the limit
            self.speed_limit = 1000
            self.index = None
            vx, vy = self.ex - self.sx, self.ey - self.sy
            self.length = norm2(vx, vy)
            self.heading = math.atan2(vy, vx)
            self.dx, self.dy = vx / self.length, vy / self.length
        else:
            self.length = self.radius * (self.ph1 - self.ph0) * self.dir
            self.sx, self.sy = self.position(0, 0)
            self.ex, self.ey = self.position(self.length, 0)

    # -- Frenet transforms ---------------------------------------------------------------------------
    def position(self, lon, lat):
        if self.kind == "S":  # straight_lane.py:53-54 ; lateral = (-dy, dx)
            return (self.sx + lon * self.dx + lat * -self.dy, self.sy + lon * self.dy + lat * self.dx)
        phi = self.dir * lon / self.radius + self.ph0  # circular_lane.py:46-50
        r = self.radius - lat * self.dir
        return (self.cx + r * math.cos(phi), self.cy + r * math.sin(phi))

    def local(self, x, y):
        if self.kind == "S":  # straight_lane.py:62-67
            ax, ay = x - self.sx, y - self.sy
            return (ax * self.dx + ay * self.dy, ax * -self.dy + ay * self.dx)
        ax, ay = x - self.cx, y - self.cy  # circular_lane.py:60-67
        phi = math.atan2(ay, ax)
        phi = self.ph0 + wrap_to_pi(phi - self.ph0)
        r = norm2(ax, ay)
        return (self.dir * (phi - self.ph0) * self.radius, self.dir * (self.radius - r))

    def heading_at(self, lon):
        if self.kind == "S":
            return self.heading
        phi = self.dir * lon / self.radius + self.ph0
        return phi + math.pi / 2 * self.dir  # circular_lane.py:52-55

    def l1_distance(self, x, y):
        s, r = self.local(x, y)  # abs_lane.py:106-112
        a, b = s - self.length, -s
        return abs(r) + (a if a > 0 else 0) + (b if b > 0 else 0)

    def precedes(self, other):
        """abs_lane.py:114-119."""
        return norm2(self.ex - other.sx, self.ey - other.sy) < 1e-1

    @property
    def start(self):
        return (self.sx, self.sy)

    @property
    def end(self):
        return (self.ex, self.ey)


def extend_straight(lane, extend_length, line_types):
    """New straight lane continuing ``lane`` (create_block_utils.py:162-170).  A clone, so it inherits
    width / speed limit / colour of the source lane like the reference's deepcopy does."""
    n = lane.clone()
    n.sx, n.sy = lane.ex, lane.ey
    n.ex, n.ey = lane.position(lane.length + extend_length, 0)
    n.line_types = line_types
    n.refresh()
    return n


def bend_then_straight(prev, follow_len, radius, angle, clockwise, width, line_types, speed_limit=20):
    """An arc tangent to the end of straight lane ``prev`` plus the straight lane leaving the arc
    (create_block_utils.py:16-59).  ``angle`` in radians."""
    bd = 1 if clockwise else -1
    center = prev.position(prev.length, bd * radius)
    x, y = -prev.dy, prev.dx  # direction_lateral
    ph0 = 0
    if y == 0:
        ph0 = 0 if x < 0 else -np.pi
    elif x == 0:
        ph0 = np.pi / 2 if y < 0 else -np.pi / 2
    else:
        base = np.arctan(y / x)
        if x < 0:
            ph0 = base
        elif y < 0:
            ph0 = np.pi + base
        elif y > 0:
            ph0 = -np.pi + base
    ph1 = ph0 + angle
    if not clockwise:
        ph0 = ph0 - np.pi
        ph1 = ph0 - angle
    bend = Lane.arc(center, radius, ph0, ph1, clockwise, width, line_types, speed_limit)
    bx, by = bend.position(2 * radius * angle / 2, 0)
    vx, vy = bx - center[0], by - center[1]
    vl = norm2(vx, vy)
    # get_vertical_vector: ((vy,-vx)/|v|, (-vy,vx)/|v|); clockwise arcs leave along the second one
    nx, ny = ((vy / vl, -vx / vl) if not clockwise else (-vy / vl, vx / vl))
    follow = Lane.straight((bx, by), (nx * follow_len + bx, ny * follow_len + by), width, line_types, speed_limit)
    return bend, follow


# ---------------------------------------------------------------------------------------------------
def neg_road(road):
    """Road.__neg__ (road.py:24-29)."""
    s, e = road
    k = e.find("-")
    if k == -1:
        return ("-" + e, "-" + s)
    return (e[k + 1:], s[k + 1:])


def is_negative(road):
    return road[1].find("-") != -1  # road.py:31-32


class RoadNet:
    def __init__(self):
        self.g = {}

    def lanes(self, road):
        return self.g[road[0]][road[1]]

    def add_lane(self, frm, to, lane):
        self.g.setdefault(frm, {}).setdefault(to, []).append(lane)

    def roads(self):
        for frm, td in self.g.items():
            for to, lanes in td.items():
                yield (frm, to), lanes

    def deco_lanes(self):
        return self.g[DECO[0]][DECO[1]] if DECO[0] in self.g else []

    def merge(self, other):
        """road_network.py:35-46: node sets must be disjoint; decoration lanes are pooled and the
        decoration entry moves to the end of the iteration order."""
        a = set(self.g) - set(DECO)
        b = set(other.g) - set(DECO)
        if a & b:
            raise ValueError("Same start node {} in two road network".format(a & b))
        deco = self.deco_lanes() + other.deco_lanes()
        self.g.update(dict(other.g))
        if deco:
            self.g.pop(DECO[0], None)
            self.g[DECO[0]] = {DECO[1]: deco}

    def subtract(self, other):
        """road_network.py:48-57."""
        for k in (self.g.keys() & other.g.keys()) - set(DECO):
            self.g.pop(k, None)
        if DECO[0] in other.g:
            mine = self.g[DECO[0]][DECO[1]]
            for lane in other.g[DECO[0]][DECO[1]]:
                if lane in mine:
                    mine.remove(lane)

    def remove_road(self, road):
        ret = self.g[road[0]].pop(road[1])
        if not self.g[road[0]]:
            self.g.pop(road[0])
        return ret

    def paths(self, start, goal):
        """Breadth-first enumeration of simple paths (road_network.py:241-256), lazily like the
        reference so that callers may remove roads between results.  Children are visited in graph
        insertion order (the reference iterates a hash-ordered set; PG maps have unique shortest paths)."""
        queue = [(start, [start])]
        while queue:
            node, path = queue.pop(0)
            if node not in self.g:
                yield []
                continue
            for nxt in [k for k in self.g[node].keys() if k not in path]:
                if nxt == goal:
                    yield path + [nxt]
                elif nxt in self.g:
                    queue.append((nxt, path + [nxt]))

    def shortest_path(self, start, goal):
        assert start != goal
        return next(self.paths(start, goal), [])

    def remove_all_roads(self, start, end):
        """road_network.py:120-133."""
        removed = []
        for path in self.paths(start, end):
            for a, b in zip(path[:-1], path[1:]):
                removed += self.remove_road((a, b))
        return removed

    def positive_lanes(self):
        """road_network.py:79-89: roads whose end node carries no '-' and that are not decoration."""
        return [lanes for road, lanes in self.roads() if not is_negative(road) and road != DECO]


# ---------------------------------------------------------------------------------------------------
def _contour(lanes, extra=3):
    """Key points bounding a road (scene_utils.py:86-128)."""
    pts = []
    first = lanes[0]
    if first.kind == "S":
        for lane, d in ((lanes[0], -1), (lanes[-1], 1)):
            pts.append(lane.position(0.1, d * (lane.width / 2.0 + extra)))
            pts.append(lane.position(lane.length - 0.1, d * (lane.width / 2.0 + extra)))
        return pts
    pi_2 = np.pi / 2.0
    for lane, d in ((lanes[0], -1), (lanes[-1], 1)):
        pts.append(lane.position(0.1, d * (lane.width / 2.0 + extra)))
        pts.append(lane.position(lane.length - 0.1, d * (lane.width / 2.0 + extra)))
        ph = (lane.ph0 // pi_2) * pi_2
        ph += pi_2 if lane.dir == 1 else 0
        for k in range(4):
            phi = ph + k * pi_2 * lane.dir
            if lane.dir * phi > lane.dir * lane.ph1:
                break
            r = lane.radius - d * (lane.width / 2.0 + extra) * lane.dir
            pts.append((lane.cx + r * math.cos(phi), lane.cy + r * math.sin(phi)))
    return pts


def road_bbox(lanes, extra=3):
    """(x_max, x_min, y_max, y_min) (scene_utils.py:74-83)."""
    pts = _contour(lanes, extra)
    xs = [p[0] for p in pts]
    ys = [p[1] for p in pts]
    return max(xs), min(xs), max(ys), min(ys)


def lane_crosses_network(net, lane, positive=0.0, ignored=None):
    """True when points sampled every metre along ``lane`` (offset ``positive * width / 2``) fall inside
    a lane already in ``net`` (scene_utils.py:40-71).  Sample points do not depend on the road being
    tested, so they are computed once."""
    x_max_2, x_min_2, y_max_2, y_min_2 = road_bbox([lane])
    samples = None
    for road, lanes in net.roads():
        if ignored and road == ignored:
            continue
        if road == DECO or not lanes:
            continue
        x_max_1, x_min_1, y_max_1, y_min_1 = road_bbox(lanes)
        if x_min_1 > x_max_2 or x_min_2 > x_max_1 or y_min_1 > y_max_2 or y_min_2 > y_max_1:
            continue
        if samples is None:
            samples = [lane.position(i, positive * lane.width / 2.0) for i in range(1, int(lane.length), 1)]
        for other in lanes:
            half = other.width / 2.0
            for (px, py) in samples:
                lon, lat = other.local(px, py)
                if math.fabs(lat) <= half and 0 <= lon <= other.length:
                    return True
    return False
