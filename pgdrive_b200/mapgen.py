"""Procedural map generation (host side): block search -> block sequence -> analytic lanes.

This restates WHAT the reference's BIG search and its eight PGDrive-v0 block types produce
(/root/reference/pgdrive/component/algorithm/BIG.py:67-151, component/blocks/*.py,
component/map/pg_map.py:34-71) so that a seed gives the same road network, but it is organised
differently: a block is a plain record, every block type is one builder function writing lanes into
that record's private ``RoadNet``, and the search is an explicit loop instead of a state machine.

The result is checked lane-by-lane against fixtures produced by the unmodified reference
(tests/golden/maps_*.json.gz, made by tools/make_golden.py).
"""
import math
from collections import OrderedDict

import numpy as np

from . import rng
from .roadnet import (
    BROKEN, CONTINUOUS, DECO, GREY, NONE, SIDE, SIDEWALK_LINE_DIST, SIDEWALK_WIDTH, YELLOW, Lane, RoadNet,
    bend_then_straight, extend_straight, lane_crosses_network, neg_road
)

# block id -> (probability in BLOCK_TYPE_DISTRIBUTION_V2, parameter space); the ORDER is the order of the
# reference's dict (blocks_prob_dist.py:31-49) because ``choice(p=...)`` indexes into it.  The six
# zero-probability types still occupy slots of the probability vector.
BLOCK_ORDER = ["C", "S", "r", "R", "X", "T", "O", "f", "F", "y", "Y", "P", "$"]
BLOCK_PROB = [0.3, 0.1, 0.1, 0.1, 0.15, 0.15, 0.1, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0]

PARAM_SPACE = {  # utils/space.py:263-306
    "I": {},
    "S": {"length": ("f", 40.0, 80.0)},
    "C": {"length": ("f", 40.0, 80.0), "radius": ("f", 25.0, 60.0), "angle": ("f", 45, 135), "dir": ("i", 0, 1)},
    "X": {"radius": ("c", 10), "change_lane_num": ("i", 0, 1), "decrease_increase": ("i", 0, 1)},
    "T": {"radius": ("c", 10), "t_type": ("i", 0, 2), "change_lane_num": ("i", 0, 1), "decrease_increase": ("i", 0, 1)},
    "O": {"exit_radius": ("f", 5, 15), "inner_radius": ("f", 15, 45), "angle": ("c", 60)},
    "r": {"length": ("f", 20, 40)},
    "R": {"length": ("f", 20, 40)},
}
SOCKET_NUM = {"I": 1, "S": 1, "C": 1, "X": 3, "T": 2, "O": 3, "r": 1, "R": 1}

MAX_TRIAL = 2  # BIG.py:28
EXIT_PART = 30  # InterSection / Roundabout EXIT_PART_LENGTH
RAMP_RADIUS, RAMP_ANGLE, RAMP_SPEED, RAMP_CONNECT, RAMP_LEN = 40, 10, 12, 20, 15
RAMP_LINES = (CONTINUOUS, CONTINUOUS)


class Socket:
    __slots__ = ("pos", "neg", "index")

    def __init__(self, pos, neg=None, index=None):
        self.pos = pos
        self.neg = neg
        self.index = index


def socket_of(road):
    return Socket(road, neg_road(road))


class Block:
    """One placed block: its lanes (``net``), the sockets later blocks may attach to, and the roads
    traffic is spawned on."""
    def __init__(self, bid, idx, pre_socket, world, seed, nocheck):
        self.id = bid
        self.idx = idx
        self.name = "%d%s" % (idx, bid)
        self.pre_socket = pre_socket
        self.pre_socket_index = pre_socket.index if pre_socket is not None else None
        self.world = world
        self.net = RoadNet()
        self.sockets = OrderedDict()
        self.respawn = []
        self.ring_spawn = []  # Roundabout.intermediate_spawn_places
        self.nocheck = nocheck
        self.trials = 0
        self.part = 0
        self.road_no = 0
        self.rs = rng.seeded(seed)
        self.params = rng.sample_space(PARAM_SPACE[bid], self.rs)  # BaseRunnable.__init__ samples once
        if idx != 0:
            self.pos_lanes = world.lanes(pre_socket.pos)
            self.neg_lanes = world.lanes(pre_socket.neg)
            self.n_pos = len(self.pos_lanes)
            self.basic = self.pos_lanes[-1]
            self.lane_width = self.basic.width

    # node naming (pg_block.py:179-201)
    def node(self, part, road):
        return "%d%s%d_%d_" % (self.idx, self.id, part, road)

    def set_part(self, part):
        self.part = part
        self.road_no = 0

    def new_node(self):
        self.road_no += 1
        return self.node(self.part, self.road_no - 1)

    def add_socket(self, sock):
        if sock.index is None:
            sock.index = "%s-socket%d" % (self.name, len(self.sockets))
        self.sockets[sock.index] = sock

    def socket(self, index):
        """get_socket: intersections and roundabouts stop spawning traffic on the arm that the next
        block is attached to (intersection.py:145-149, roundabout.py:193-197)."""
        if isinstance(index, (int, np.integer)):
            index = list(self.sockets)[index]
        sock = self.sockets[index]
        if self.id in ("X", "T", "O") and sock.neg in self.respawn:
            self.respawn.remove(sock.neg)
        return sock

    def clear(self):
        self.world.subtract(self.net)
        self.net.g.clear()
        self.part = 0
        self.road_no = 0
        self.respawn = []
        self.sockets.clear()

    def build(self, config=None):
        """construct_block (base_block.py:72-96): resample parameters, rebuild topology, merge."""
        self.params = rng.sample_space(PARAM_SPACE[self.id], self.rs)
        if config:
            self.params.update(config)
        self.clear()
        self.trials += 1
        ok = BUILDERS[self.id](self)
        self.world.merge(self.net)
        return ok

    def respawn_lanes(self):
        return [self.net.lanes(r) for r in self.respawn]

    def spawn_lanes(self):
        """get_intermediate_spawn_lanes of each block type."""
        if self.id in ("X", "T"):
            return self.respawn_lanes()
        if self.id == "O":
            return self.respawn_lanes() + self.ring_spawn
        out = self.net.positive_lanes()
        for lanes in self.respawn_lanes():
            if not any(lanes is x for x in out) and lanes not in out:
                out.append(lanes)
        return out


# ---------------------------------------------------------------------------------------------------
def _crosses(b, lane, positive, ignored=None):
    """check_lane_on_road: when checking is disabled the reference reports 'crossing' (scene_utils.py:49-50)."""
    if b.nocheck:
        return True
    return lane_crosses_network(b.world, lane, positive, ignored)


def road_from(
    b, lane, lane_num, road, toward_smaller=True, ignore=None, center=CONTINUOUS, one_side=True, side=SIDE,
    inner=BROKEN, center_color=YELLOW
):
    """Lay ``lane_num`` parallel lanes starting from ``lane`` (which becomes the outermost, or with
    ``toward_smaller=False`` the innermost, lane) and file them under ``road`` in the block's net.
    Returns False when the new road overlaps the existing world (create_block_utils.py:62-159)."""
    extra = lane_num - 1
    origin = lane
    lanes = []
    w = lane.width
    for i in range(extra, 0, -1):
        s = lane.clone()
        if lane.kind == "S":
            off = -w if toward_smaller else w
            # the copy keeps direction / length of the source; only the end points move
            s.sx, s.sy = lane.position(0, off)
            s.ex, s.ey = lane.position(lane.length, off)
        else:
            cw = lane.dir == 1
            if not toward_smaller:
                s.radius = lane.radius - w if cw else lane.radius + w
            else:
                s.radius = lane.radius + w if cw else lane.radius - w
            s.refresh()
        if i == 1:
            s.line_types = [center, inner] if toward_smaller else [inner, side]
        else:
            s.line_types = [inner, inner]
        lanes.append(s)
        lane = s
    if toward_smaller:
        lanes.reverse()
        lanes.append(origin)
        origin.line_types = [inner if len(lanes) > 1 else center, side]
    else:
        lanes.insert(0, origin)
        if len(lanes) > 1:
            origin.line_types = [origin.line_types[0], lanes[-1].line_types[0]]
    factor = (SIDEWALK_WIDTH + SIDEWALK_LINE_DIST + w / 2.0) * 2.0 / w
    if one_side:
        ok = not _crosses(b, origin, factor, ignore)
    else:
        ok = not (_crosses(b, origin, factor, ignore) or _crosses(b, lanes[0], -0.95, ignore))
    for ln in lanes:
        b.net.add_lane(road[0], road[1], ln)
    if extra == 0:
        lanes[-1].line_types = [center, side]
    lanes[0].line_color = [center_color, GREY]
    return ok


def adverse_road(b, road, ignore=None, center=CONTINUOUS, side=SIDE, inner=BROKEN, center_color=YELLOW):
    """Mirror ``road`` across its centre line to make the opposite carriageway
    (create_block_utils.py:177-230)."""
    lanes = b.net.lanes(road)
    ref = lanes[-1]
    num = len(lanes) * 2
    w = ref.width
    if ref.kind == "S":
        start = ref.position(ref.length, -(num - 1) * w)
        end = ref.position(0, -(num - 1) * w)
        sym = Lane.straight(start, end, w, ref.line_types, ref.speed_limit)
    else:
        cw = ref.dir != 1
        radius = ref.radius + (num - 1) * w if not cw else ref.radius - (num - 1) * w
        sym = Lane.arc((ref.cx, ref.cy), radius, ref.ph1, ref.ph0, cw, w, ref.line_types, ref.speed_limit)
    ok = road_from(b, sym, num // 2, neg_road(road), ignore=ignore, side=side, inner=inner, center=center,
                   center_color=center_color)
    b.net.lanes(road)[0].line_color = [center_color, GREY]
    return ok


# ---------------------------------------------------------------------------------------------------
def build_first(world, lane_width, lane_num, length=50, nocheck=False):
    """FirstPGBlock (first_block.py:25-89): 10 m entrance road + (length-10) m exit road, both ways."""
    b = Block("I", 0, None, world, 0, nocheck)
    b.pre_socket = Socket(DECO, DECO)
    basic = Lane.straight((0, lane_width * (lane_num - 1)), (10, lane_width * (lane_num - 1)), lane_width,
                          (BROKEN, SIDE))
    r1 = (">", ">>")
    road_from(b, basic, lane_num, r1)
    adverse_road(b, r1)
    nxt = extend_straight(basic, length - 10, [BROKEN, SIDE])
    r2 = (">>", ">>>")
    road_from(b, nxt, lane_num, r2)
    adverse_road(b, r2)
    world.merge(b.net)
    sock = socket_of(r2)
    sock.index = "0I-socket0"
    b.add_socket(sock)
    b.respawn = [r2]
    return b


def _straight(b):
    b.set_part(0)
    new = extend_straight(b.basic, b.params["length"], [BROKEN, SIDE])
    road = (b.pre_socket.pos[1], b.new_node())
    ok = road_from(b, new, b.n_pos, road)
    ok = adverse_road(b, road) and ok
    b.add_socket(socket_of(road))
    return ok


def _curve(b):
    p = b.params
    road = (b.pre_socket.pos[1], b.new_node())
    bend, straight = bend_then_straight(
        b.basic, p["length"], p["radius"], np.deg2rad(p["angle"]), p["dir"], b.basic.width, (BROKEN, SIDE)
    )
    ok = road_from(b, bend, b.n_pos, road)
    ok = adverse_road(b, road) and ok
    road = (road[1], b.new_node())
    ok = road_from(b, straight, b.n_pos, road) and ok
    ok = adverse_road(b, road) and ok
    b.add_socket(socket_of(road))
    return ok


def _intersection(b):
    """InterSection._try_plug_into_previous_block (intersection.py:45-96); Std variants force
    change_lane_num = 0 (std_intersection.py:6-9)."""
    p = b.params
    p["change_lane_num"] = 0
    di = -1 if p["decrease_increase"] == 0 else 1
    if b.n_pos <= 1:
        di = 1
    elif b.n_pos >= 4:
        di = -1
    b.n_cross = b.n_pos + di * p["change_lane_num"]
    ok = True
    attach = b.pre_socket.pos
    attach_lanes = b.world.lanes(attach)
    nodes = [b.node(0, 0), b.node(1, 0), b.node(2, 0), b.pre_socket.neg[0]]
    for i in range(4):
        right_lane, good = _intersection_part(b, attach_lanes, attach, p["radius"], nodes, i)
        nodes = nodes[1:] + nodes[:1]
        ok = ok and good
        if i != 3:
            n = b.n_pos if i == 1 else b.n_cross
            exit_road = (b.node(i, 0), b.node(i, 1))
            ok = road_from(b, right_lane, n, exit_road) and ok
            ok = adverse_road(b, exit_road) and ok
            sock = socket_of(exit_road)
            b.respawn.append(sock.neg)
            b.add_socket(sock)
            attach = sock.neg
            attach_lanes = b.net.lanes(attach)
    return ok


def _intersection_part(b, attach_lanes, attach, radius, nodes, part):
    n = b.n_cross if part in (0, 2) else b.n_pos
    left = attach_lanes[0]
    w = left.width
    n_turn = min(b.n_pos, b.n_cross)
    # left turn (intersection.py:151-206)
    left_r = radius + n * w
    diff = b.n_cross - b.n_pos
    if (part in (1, 3) and diff > 0) or (part in (0, 2) and diff < 0):
        diff = abs(diff)
        bend, extra = bend_then_straight(left, b.lane_width * diff, left_r, np.deg2rad(90), False, w, (NONE, NONE))
        mid = nodes[2] + "extra"
        road_from(b, bend, n_turn, (attach[1], mid), toward_smaller=False, center=NONE, side=NONE, inner=NONE)
        road_from(b, extra, n_turn, (mid, nodes[2]), toward_smaller=False, center=NONE, side=NONE, inner=NONE)
    else:
        bend, _ = bend_then_straight(left, EXIT_PART, left_r, np.deg2rad(90), False, w, (NONE, NONE))
        road_from(b, bend, n_turn, (attach[1], nodes[2]), toward_smaller=False, center=NONE, side=NONE, inner=NONE)
    # straight through
    src = [ln.clone() for ln in attach_lanes]
    through = 2 * radius + (2 * n - 1) * src[0].width
    for ln in src:
        b.net.add_lane(attach[1], nodes[1], extend_straight(ln, through, (NONE, NONE)))
    # right turn
    right = src[-1]
    rbend, rstraight = bend_then_straight(right, EXIT_PART, radius, np.deg2rad(90), True, right.width, (NONE, SIDE))
    ok = not _crosses(b, rbend, 1)
    road_from(b, rbend, n_turn, (attach[1], nodes[0]), toward_smaller=True, side=SIDE, inner=NONE, center=NONE)
    rstraight.line_types = [BROKEN, SIDE]
    return rstraight, ok


def _t_intersection(b):
    """TInterSection (t_intersection.py:17-86): build the 4-arm crossing, then delete one arm."""
    ok = _intersection(b)
    t = b.params["t_type"]
    pre = b.pre_socket
    b.add_socket(pre)
    gone = b.sockets["%s-socket%d" % (b.name, t)]
    start_node, end_node = gone.neg[1], gone.pos[0]
    for i in range(4):
        if i == t:
            continue
        s = b.sockets["%s-socket%d" % (b.name, i)] if i < 3 else b.sockets[pre.index]
        exit_node = s.pos[0] if i != 3 else s.neg[0]
        b.net.remove_all_roads(start_node, exit_node)
        entry_node = s.neg[1] if i != 3 else s.pos[1]
        b.net.remove_all_roads(entry_node, end_node)
    _t_relabel(b, t)
    b.sockets.pop(pre.index)
    sock = b.sockets.pop("%s-socket%d" % (b.name, t))
    b.net.remove_all_roads(sock.pos[0], sock.pos[1])
    b.net.remove_all_roads(sock.neg[0], sock.neg[1])
    b.respawn.remove(sock.neg)
    return ok


def _t_relabel(b, t):
    """_change_vis (t_intersection.py:22-51): the through road opposite the removed arm gets real
    lane lines (this changes which lines end an episode, so it is not cosmetic here)."""
    socks = list(b.sockets.values())
    nxt = socks[(t + 1) % 4]
    last = socks[(t + 3) % 4]
    n_pos, n_neg = nxt.pos, nxt.neg
    l_pos, l_neg = last.pos, last.neg
    if t == 2:  # Goal.LEFT
        n_pos, n_neg = nxt.neg, nxt.pos
    if t == 0:  # Goal.RIGHT
        l_pos, l_neg = last.neg, last.pos
    for i, road in enumerate([(l_neg[1], n_pos[0]), (n_neg[1], l_pos[0])]):
        lanes = b.net.lanes(road)
        outside = SIDE if i == 0 else NONE
        for k, lane in enumerate(lanes):
            lane.line_types = [BROKEN, BROKEN] if k != len(lanes) - 1 else [BROKEN, outside]
            if k == 0:
                lane.line_color = [YELLOW, GREY]
                if i == 1:
                    lane.line_types[0] = NONE


def _roundabout(b):
    b.ring_spawn = []
    p = b.params
    ok = True
    attach = b.pre_socket.pos
    for i in range(4):
        exit_road, good = _roundabout_part(b, attach, i, p["exit_radius"], p["inner_radius"], p["angle"])
        ok = ok and good
        if i < 3:
            ok = adverse_road(b, exit_road) and ok
            attach = neg_road(exit_road)
    b.respawn += [s.neg for s in b.sockets.values()]
    return ok


def _tool_lane(straight, back):
    return Lane.straight(straight.position(-back, 0), straight.position(0, 0), Lane.DEFAULT_WIDTH)


def _roundabout_part(b, road, part, r_exit, r_inner, angle):
    """One quarter of the ring (roundabout.py:49-191)."""
    ok = True
    b.set_part(part)
    n = b.n_pos
    w = b.lane_width
    r_big = (n * 2 - 1) * w + r_inner
    # entry arc
    seg = (road[1], b.new_node())
    lanes = b.world.lanes(road) if part == 0 else b.net.lanes(road)
    bend, straight = bend_then_straight(lanes[-1], 10, r_exit, np.deg2rad(angle), True, w, (BROKEN, SIDE))
    skip = (b.node((part + 3) % 4, 0), b.node((part + 3) % 4, 0))
    ok = road_from(b, bend, n, seg, ignore=skip) and ok
    for k, ln in enumerate(b.net.lanes(seg)):
        ln.line_types = [NONE, SIDE] if k == n - 1 else [NONE, NONE]
    # ring arc
    bend, to_next = bend_then_straight(
        _tool_lane(straight, 5), 10, r_big, np.deg2rad(2 * angle - 90), False, w, (BROKEN, SIDE)
    )
    seg = (seg[1], b.new_node())
    ok = road_from(b, bend, n, seg) and ok
    b.ring_spawn.append(b.net.lanes(seg))
    # exit arc + exit straight
    bend, straight = bend_then_straight(_tool_lane(to_next, 5), EXIT_PART, r_exit, np.deg2rad(angle), True, w,
                                        (BROKEN, SIDE))
    seg = (seg[1], b.new_node() if part < 3 else b.pre_socket.neg[0])
    ok = road_from(b, bend, n, seg) and ok
    for k, ln in enumerate(b.net.lanes(seg)):
        ln.line_types = [NONE, SIDE] if k == n - 1 else [NONE, NONE]
    exit_road = (seg[1], b.new_node())
    if part < 3:
        ok = road_from(b, straight, n, exit_road) and ok
        b.add_socket(socket_of(exit_road))
    # inner connector to the next quarter
    seg = (b.node(part, 1), b.node((part + 1) % 4, 0))
    beneath = (n * 2 - 1) * w / 2 + r_exit
    r_seg = beneath / math.cos(np.deg2rad(angle)) - r_exit
    bend, _ = bend_then_straight(_tool_lane(to_next, 6), 5, r_seg, np.deg2rad(180 - 2 * angle), False, w,
                                 (BROKEN, SIDE))
    road_from(b, bend, n, seg)
    for k, ln in enumerate(b.net.lanes(seg)):
        if k == 0:
            ln.line_types = [CONTINUOUS, BROKEN] if n > 1 else [CONTINUOUS, NONE]
        else:
            ln.line_types = [BROKEN, BROKEN]
    return exit_road, ok


def _in_ramp(b):
    """InRampOnStraight (ramp.py:43-204)."""
    acc_len = b.params["length"]
    n, w = b.n_pos, b.lane_width
    extra_part, socket_len = 10, 20
    ok = True
    b.set_part(0)
    sin_a, cos_a = math.sin(np.deg2rad(RAMP_ANGLE)), math.cos(np.deg2rad(RAMP_ANGLE))
    lon_len = sin_a * RAMP_RADIUS * 2 + cos_a * RAMP_CONNECT + RAMP_LEN
    extend = extend_straight(b.basic, lon_len + extra_part, [BROKEN, CONTINUOUS])
    extend_road = (b.pre_socket.pos[1], b.new_node())
    ok = road_from(b, extend, n, extend_road, side=CONTINUOUS) and ok
    b.net.lanes(extend_road)[-1].line_types = [BROKEN if n != 1 else CONTINUOUS, CONTINUOUS]
    ok = adverse_road(b, extend_road) and ok
    b.net.lanes(neg_road(extend_road))[-1].line_types = [NONE if n == 1 else BROKEN, SIDE]
    # acceleration part
    acc_side = extend_straight(extend, acc_len + w, [extend.line_types[0], SIDE])
    acc_road = (extend_road[1], b.new_node())
    ok = road_from(b, acc_side, n, acc_road, side=CONTINUOUS) and ok
    ok = adverse_road(b, acc_road) and ok
    b.net.lanes(acc_road)[-1].line_types = [CONTINUOUS if n == 1 else BROKEN, BROKEN]
    # socket part
    sock_side = extend_straight(acc_side, socket_len, acc_side.line_types)
    sock_road = (acc_road[1], b.new_node())
    ok = road_from(b, sock_side, n, sock_road, side=CONTINUOUS) and ok
    ok = adverse_road(b, sock_road) and ok
    b.add_socket(socket_of(sock_road))
    # the ramp itself
    b.set_part(1)
    lat = (1 - cos_a) * RAMP_RADIUS * 2 + sin_a * RAMP_CONNECT
    end_pt = extend.position(extra_part + RAMP_LEN, lat + w)
    start_pt = extend.position(extra_part, lat + w)
    entry = Lane.straight(start_pt, end_pt, w, RAMP_LINES, RAMP_SPEED)
    entry_road = (b.new_node(), b.new_node())
    b.net.add_lane(entry_road[0], entry_road[1], entry)
    ok = (not _crosses(b, entry, 0.95)) and ok
    b.respawn.append(entry_road)
    bend1, connect = bend_then_straight(entry, RAMP_CONNECT, RAMP_RADIUS, np.deg2rad(RAMP_ANGLE), False, w,
                                        RAMP_LINES, RAMP_SPEED)
    bend1_road = (entry_road[1], b.new_node())
    connect_road = (bend1_road[1], b.new_node())
    b.net.add_lane(bend1_road[0], bend1_road[1], bend1)
    b.net.add_lane(connect_road[0], connect_road[1], connect)
    ok = (not _crosses(b, bend1, 0.95)) and ok
    ok = (not _crosses(b, connect, 0.95)) and ok
    bend2, acc_lane = bend_then_straight(connect, acc_len, RAMP_RADIUS, np.deg2rad(RAMP_ANGLE), True, w,
                                         RAMP_LINES, RAMP_SPEED)
    acc_lane.line_types = [BROKEN, CONTINUOUS]
    bend2_road = (connect_road[1], b.node(0, 0))
    b.net.add_lane(bend2_road[0], bend2_road[1], bend2)
    b.net.add_lane(acc_road[0], acc_road[1], acc_lane)
    ok = (not _crosses(b, bend2, 0.95)) and ok
    ok = (not _crosses(b, acc_lane, 0.95)) and ok
    merge, _ = bend_then_straight(acc_lane, 10, w / 2, np.pi / 2, False, w, (BROKEN, CONTINUOUS))
    b.net.add_lane(DECO[0], DECO[1], merge)
    return ok


def _out_ramp(b):
    """OutRampOnStraight (ramp.py:235-365)."""
    n, w = b.n_pos, b.lane_width
    ok = True
    sin_a, cos_a = math.sin(np.deg2rad(RAMP_ANGLE)), math.cos(np.deg2rad(RAMP_ANGLE))
    lon_len = sin_a * RAMP_RADIUS * 2 + cos_a * RAMP_CONNECT + RAMP_LEN + 15
    b.set_part(0)
    dec_len = b.params["length"]
    dec_lane = extend_straight(b.basic, dec_len + w, [b.basic.line_types[0], SIDE])
    dec_road = (b.pre_socket.pos[1], b.new_node())
    ok = road_from(b, dec_lane, n, dec_road, side=CONTINUOUS) and ok
    ok = adverse_road(b, dec_road) and ok
    dec_right = b.net.lanes(dec_road)[-1]
    dec_right.line_types = [CONTINUOUS if n == 1 else BROKEN, NONE]
    extend = extend_straight(dec_right, lon_len, [dec_right.line_types[0], CONTINUOUS])
    extend_road = (dec_road[1], b.new_node())
    ok = road_from(b, extend, n, extend_road, side=CONTINUOUS) and ok
    ok = adverse_road(b, extend_road) and ok
    b.net.lanes(neg_road(extend_road))[-1].line_types = [NONE if n == 1 else BROKEN, SIDE]
    b.add_socket(socket_of(extend_road))
    # deceleration lane + ramp
    b.set_part(1)
    side_lane = Lane.straight(dec_right.position(w, w), dec_right.position(dec_right.length, w), w,
                              (BROKEN, CONTINUOUS))
    b.net.add_lane(dec_road[0], dec_road[1], side_lane)
    ok = (not _crosses(b, side_lane, 0.95)) and ok
    bend1, connect = bend_then_straight(side_lane, RAMP_CONNECT, RAMP_RADIUS, np.deg2rad(RAMP_ANGLE), True, w,
                                        RAMP_LINES, RAMP_SPEED)
    bend1_road = (dec_road[1], b.new_node())
    connect_road = (bend1_road[1], b.new_node())
    b.net.add_lane(bend1_road[0], bend1_road[1], bend1)
    b.net.add_lane(connect_road[0], connect_road[1], connect)
    ok = (not _crosses(b, bend1, 0.95)) and ok
    ok = (not _crosses(b, connect, 0.95)) and ok
    bend2, exit_lane = bend_then_straight(connect, RAMP_LEN, RAMP_RADIUS, np.deg2rad(RAMP_ANGLE), False, w,
                                          RAMP_LINES, RAMP_SPEED)
    bend2_road = (connect_road[1], b.new_node())
    exit_road = (bend2_road[1], b.new_node())
    b.net.add_lane(bend2_road[0], bend2_road[1], bend2)
    b.net.add_lane(exit_road[0], exit_road[1], exit_lane)
    ok = (not _crosses(b, bend2, 0.95)) and ok
    ok = (not _crosses(b, exit_lane, 0.95)) and ok
    tool = Lane.straight(side_lane.end, side_lane.start, side_lane.width)
    deco, _ = bend_then_straight(tool, 10, w / 2, np.pi / 2, True, w, (CONTINUOUS, BROKEN))
    b.net.add_lane(DECO[0], DECO[1], deco)
    return ok


BUILDERS = {"S": _straight, "C": _curve, "X": _intersection, "T": _t_intersection, "O": _roundabout,
            "r": _in_ramp, "R": _out_ramp}


# ---------------------------------------------------------------------------------------------------
class PGMapData:
    """A generated map: the merged road network, the placed blocks and the serialisable block
    sequence (the reference's ``save_map`` format, base_map.py:103-118)."""
    def __init__(self, seed, net, blocks, sequence, lane_num, lane_width):
        self.seed = seed
        self.net = net
        self.blocks = blocks
        self.block_sequence = sequence
        self.lane_num = lane_num
        self.lane_width = lane_width

    @property
    def num_blocks(self):
        return len(self.blocks)

    @property
    def road_network(self):
        return self.net

    def save_map(self):
        """BaseMap.save_map (component/map/base_map.py:103-118)."""
        return dict(block_sequence=[dict(b) for b in self.block_sequence])


def search_sequence(seed, block_num=3, lane_num=3, lane_width=3.5, exit_length=50, sequence=None):
    """The BIG search: append random blocks, retry a block up to MAX_TRIAL times when it overlaps the
    map so far, then back-track (BIG.py:67-151).  Returns the block sequence in save_map form."""
    rs = rng.seeded(seed)
    world = RoadNet()
    blocks = [build_first(world, lane_width, lane_num, exit_length)]
    target = (block_num if sequence is None else len(sequence)) + 1
    FORWARD, DESTRUCT, SIBLING, BACK = range(4)
    step = FORWARD
    while not (len(blocks) >= target and step == FORWARD):
        if step == FORWARD:
            if sequence is None:
                bid = BLOCK_ORDER[int(rs.choice(len(BLOCK_ORDER), p=BLOCK_PROB))]
            else:
                bid = sequence[len(blocks) - 1]
            prev = blocks[-1]
            sock_idx = str(rs.choice(list(prev.sockets)))
            blk = Block(bid, len(blocks), prev.socket(sock_idx), world, int(rs.randint(0, 10000)), False)
            blocks.append(blk)
            step = FORWARD if blk.build() else DESTRUCT
        elif step == DESTRUCT:
            blk = blocks[-1]
            blk.clear()
            step = SIBLING if blk.trials < MAX_TRIAL else BACK
        elif step == SIBLING:
            blk = blocks[-1]
            if blk.trials < MAX_TRIAL:
                step = FORWARD if blk.build() else DESTRUCT
            else:
                step = BACK
        else:  # BACK
            blocks.pop()
            blocks[-1].clear()
            step = SIBLING
    seq = []
    for blk in blocks:
        rec = dict(blk.params)
        rec["id"] = blk.id
        rec["pre_block_socket_index"] = blk.pre_socket_index
        seq.append(rec)
    return seq


def build_from_sequence(seed, sequence, lane_num=3, lane_width=3.5, exit_length=50):
    """_config_generate (pg_map.py:48-71): the map the simulator actually drives on is rebuilt from the
    stored block sequence with overlap checking off."""
    world = RoadNet()
    blocks = [build_first(world, lane_width, lane_num, exit_length, nocheck=True)]
    for i, rec in enumerate(sequence[1:], 1):
        cfg = {k: v for k, v in rec.items() if k not in ("id", "pre_block_socket_index")}
        prev = blocks[-1]
        blk = Block(rec["id"], i, prev.socket(rec["pre_block_socket_index"]), world, seed, True)
        blk.build(cfg)
        blocks.append(blk)
    return PGMapData(seed, world, blocks, sequence, lane_num, lane_width)


def generate_map(seed, block_num=3, lane_num=3, lane_width=3.5, exit_length=50, sequence=None):
    seq = search_sequence(seed, block_num, lane_num, lane_width, exit_length, sequence)
    return build_from_sequence(seed, seq, lane_num, lane_width, exit_length)
