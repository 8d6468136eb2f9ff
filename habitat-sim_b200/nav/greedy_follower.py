"""GreedyGeodesicFollower on top of the batched PathFinder, for one agent or N agents at once.

Restates esp::nav::GreedyGeodesicFollowerImpl (src/esp/nav/GreedyFollower.cpp:38-240) and the
Python wrapper src_python/habitat_sim/nav/greedy_geodesic_follower.py:18-200.  The reference
evaluates each primitive `[LEFT]*n + [FORWARD]` / `[RIGHT]*n + [FORWARD]` by bouncing
C++ -> Python -> C++ through a MoveFn (GreedyFollower.h:54) and running try_step, find_path
and distance_to_closest_obstacle one at a time (GreedyFollower.cpp:46-81).  Here the primitives
of ALL agents of a decision (N agents x 2*ceil(pi/turn) headings) are evaluated by ONE batched
try_steps / geodesic_distances / distances_to_closest_obstacle call each, and
the reference's sequential selection (first strictly better reward, early exit above 0.99,
GreedyFollower.cpp:99-137) is replayed over the results, vectorised over the agents.  The
single-agent classes are the batch of one, so both give the same action sequences.

The forward / turn kinematics are those of habitat_sim/agent/controls/default_controls.py
(move along local -Z, rotate about +Y) followed by ObjectControls.action's step filter and
collision test (object_controls.py:50-90).  Rotations are quaternions (x, y, z, w), float64.

`pathfinder` is anything with try_steps(starts, ends, allow_sliding), geodesic_distances(starts,
ends) and distances_to_closest_obstacle(pts, radius) over [N,3] float32 arrays.
"""
from __future__ import annotations

import enum
import math

import numpy as np


class GreedyFollowerCodes(enum.IntEnum):
    """GreedyFollower.h:39-45"""
    ERROR = -2
    STOP = -1
    FORWARD = 0
    LEFT = 1
    RIGHT = 2


_EPS = 1e-5  # object_controls.py: EPS
_CLOSE_TO_OBS = 0.2  # closeToObsThreshold_, GreedyFollower.h:129
_COLLISION_COST = 0.25  # collisionCost_, GreedyFollower.h:130


# ---- quaternion kinematics over [M,4] / [M,3] float64 arrays (explicit component arithmetic, so
# ---- that a batch of one and a batch of many round identically) ---------------------------------
def _quat_mul(a, b):
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw,
                     aw * bw - ax * bx - ay * by - az * bz], -1)


def _cross(u, v):
    return np.stack([u[..., 1] * v[..., 2] - u[..., 2] * v[..., 1],
                     u[..., 2] * v[..., 0] - u[..., 0] * v[..., 2],
                     u[..., 0] * v[..., 1] - u[..., 1] * v[..., 0]], -1)


def _quat_rotate(q, v):
    u = q[..., :3]
    w = q[..., 3:4]
    return v + 2.0 * _cross(u, _cross(u, v) + w * v)


def _yaw_quat(theta):
    return np.array([0.0, math.sin(theta / 2.0), 0.0, math.cos(theta / 2.0)], np.float64)


def _normalized(q):
    n = np.sqrt(q[..., 0] * q[..., 0] + q[..., 1] * q[..., 1] + q[..., 2] * q[..., 2] + q[..., 3] * q[..., 3])
    return q / n[..., None]


def _forward_target_of(follower, rot, pos):
    """Where `follower`'s FORWARD aims from (rot, pos) (used by the tests to replay actions)."""
    return follower._forward_target(np.asarray(rot, np.float64), np.asarray(pos, np.float64))


class GreedyGeodesicFollowerBatchImpl:
    """N independent GreedyGeodesicFollowerImpl state machines advanced together."""

    def __init__(self, pathfinder, num_agents: int, goal_dist: float = 0.1875, forward_amount: float = 0.25,
                 turn_amount: float = math.radians(10.0), fix_thrashing: bool = True,
                 thrashing_threshold: int = 16, allow_sliding: bool = True):
        self.pathfinder = pathfinder
        self.n = int(num_agents)
        self.goal_dist = float(goal_dist)
        self.forward_amount = float(forward_amount)
        self.turn_amount = float(turn_amount)
        self.fix_thrashing = bool(fix_thrashing)
        self.thrashing_threshold = int(thrashing_threshold)
        self.allow_sliding = bool(allow_sliding)
        # candidate headings in the reference's loop order: angle = 0, t, 2t, ... < pi (f32 accumulate)
        n_steps = 0
        angle = np.float32(0.0)
        while angle < np.float32(math.pi):
            n_steps += 1
            angle = np.float32(angle + np.float32(self.turn_amount))
        self._n_steps = n_steps
        self.reset()

    def reset(self, agents=None):
        if agents is None:
            self._actions = [[] for _ in range(self.n)]
            self._thrashing = [[] for _ in range(self.n)]
        else:
            for i in agents:
                self._actions[i] = []
                self._thrashing[i] = []

    # -- kinematics ---------------------------------------------------------------------
    def _forward_target(self, rot, pos):
        return pos + _quat_rotate(rot, np.array([0.0, 0.0, -self.forward_amount]))

    def _turn(self, rot, sign):
        return _normalized(_quat_mul(rot, _yaw_quat(sign * self.turn_amount)))

    def _geo(self, pos, end):
        return np.asarray(self.pathfinder.geodesic_distances(np.asarray(pos, np.float32).reshape(-1, 3),
                                                             np.asarray(end, np.float32).reshape(-1, 3)),
                          np.float32)

    # -- GreedyFollower.cpp:83-140, for the agents `rot`/`pos`/`end`/`geo` describe -------------
    def _next_best_prims(self, rot, pos, end, geo):
        m = len(pos)
        out = [None] * m
        search = []
        for i in range(m):
            if geo[i] == np.inf or np.isnan(geo[i]):
                out[i] = [GreedyFollowerCodes.ERROR]
            elif geo[i] < self.goal_dist:
                out[i] = [GreedyFollowerCodes.STOP]
            else:
                search.append(i)
        if not search:
            return out
        idx = np.asarray(search)
        rot, pos, end, geo = rot[idx], pos[idx], end[idx], geo[idx]
        ms, ns = len(idx), self._n_steps
        lrot, rrot = rot.copy(), rot.copy()
        targets = np.empty((ms, 2 * ns, 3), np.float64)
        for k in range(ns):
            targets[:, 2 * k] = self._forward_target(lrot, pos)
            targets[:, 2 * k + 1] = self._forward_target(rrot, pos)
            lrot = self._turn(lrot, +1.0)
            rrot = self._turn(rrot, -1.0)
        targets = targets.astype(np.float32).reshape(-1, 3)
        starts = np.repeat(pos.astype(np.float32), 2 * ns, 0)
        ends = np.repeat(end.astype(np.float32), 2 * ns, 0)
        # one batched evaluation of every primitive of every agent (GreedyFollower.cpp:46-60)
        filt = np.asarray(self.pathfinder.try_steps(starts, targets, self.allow_sliding), np.float32)
        before = ((targets - starts).astype(np.float32) ** 2).sum(1)
        after = ((filt - starts).astype(np.float32) ** 2).sum(1)
        collided = ((after + _EPS) < before).reshape(ms, 2 * ns)
        geo_after = np.asarray(self.pathfinder.geodesic_distances(filt, ends), np.float32).reshape(ms, 2 * ns)
        obs_after = np.asarray(self.pathfinder.distances_to_closest_obstacle(filt, 1.1 * _CLOSE_TO_OBS),
                               np.float32).reshape(ms, 2 * ns)
        k_of = np.repeat(np.arange(ns), 2)[None, :]
        # computeReward, GreedyFollower.cpp:62-81, in its float arithmetic
        f32 = np.float32
        penalty = f32(-0.0125) * k_of.astype(f32)
        penalty = penalty - np.where(collided, f32(_COLLISION_COST), f32(0.0))
        penalty = penalty - np.where(obs_after < f32(_CLOSE_TO_OBS), f32(0.05), f32(0.0))
        reward = (geo.astype(f32)[:, None] - geo_after) / f32(self.forward_amount) + penalty
        best_reward = np.full(ms, -_COLLISION_COST, np.float64)
        best = np.full(ms, -1, np.int64)
        active = np.ones(ms, bool)
        for c in range(2 * ns):
            with np.errstate(invalid="ignore"):
                upd = active & (reward[:, c] > best_reward)
            best_reward[upd] = reward[upd, c]
            best[upd] = c
            if c & 1:  # RIGHT candidate: GreedyFollower.cpp:131-135 (float bestReward > 0.99f)
                active &= ~(best_reward.astype(np.float32) > np.float32(0.99))
                if not active.any():
                    break
        for j, i in enumerate(search):
            if best[j] < 0:
                out[i] = []
            else:
                side = GreedyFollowerCodes.RIGHT if (best[j] & 1) else GreedyFollowerCodes.LEFT
                out[i] = [side] * int(best[j] >> 1) + [GreedyFollowerCodes.FORWARD]
        return out

    def _decide(self, rot, pos, end):
        """Best primitive of every agent described by rot / pos / end (nextBestPrimAlong,
        GreedyFollower.cpp:83-140).  A PathFinder of this package does all of it on the device
        (hbn_follower_best_prims: kinematics, the three batched queries, rewards, selection); any
        other backend (the tests' oracle adapter) gets the numpy restatement above."""
        dev = getattr(self.pathfinder, "follower_best_prims", None)
        if dev is None:
            return self._next_best_prims(rot, pos, end, self._geo(pos, end))
        prim, _geo = dev(rot, pos, end, self.goal_dist, self.forward_amount, self.turn_amount, self._n_steps,
                         self.allow_sliding)
        out = []
        for c in prim:
            c = int(c)
            if c == -2:
                out.append([GreedyFollowerCodes.ERROR])
            elif c == -1:
                out.append([GreedyFollowerCodes.STOP])
            elif c < 0:
                out.append([])
            else:
                side = GreedyFollowerCodes.RIGHT if (c & 1) else GreedyFollowerCodes.LEFT
                out.append([side] * (c >> 1) + [GreedyFollowerCodes.FORWARD])
        return out

    def _is_thrashing(self, actions):
        if len(actions) < self.thrashing_threshold:
            return False
        last = actions[-1]
        thrashing = last in (GreedyFollowerCodes.LEFT, GreedyFollowerCodes.RIGHT)
        i = 2
        while i < self.thrashing_threshold + 1 and thrashing:
            a = actions[-i]
            thrashing = ((a == GreedyFollowerCodes.RIGHT and last == GreedyFollowerCodes.LEFT)
                         or (a == GreedyFollowerCodes.LEFT and last == GreedyFollowerCodes.RIGHT))
            last = a
            i += 1
        return thrashing

    # -- GreedyFollower.cpp:160-188 ----------------------------------------------------------
    def next_actions_along(self, rots, poss, ends):
        """One decision for every agent: [N] GreedyFollowerCodes."""
        rot = np.asarray(rots, np.float64).reshape(self.n, 4)
        pos = np.asarray(poss, np.float64).reshape(self.n, 3)
        end = np.asarray(ends, np.float64).reshape(self.n, 3)
        need = [i for i in range(self.n) if not (self.fix_thrashing and self._thrashing[i])]
        prims = {}
        if need:
            ix = np.asarray(need)
            for i, acts in zip(need, self._decide(rot[ix], pos[ix], end[ix])):
                prims[i] = acts
        out = []
        for i in range(self.n):
            if i not in prims:
                nxt = self._thrashing[i].pop()
            else:
                acts = prims[i]
                if not acts:
                    nxt = GreedyFollowerCodes.ERROR
                elif self.fix_thrashing and self._is_thrashing(self._actions[i]):
                    self._thrashing[i] = list(reversed(acts))
                    nxt = self._thrashing[i].pop()
                else:
                    nxt = acts[0]
            self._actions[i].append(nxt)
            out.append(nxt)
        return out

    # -- GreedyFollower.cpp:190-240 ----------------------------------------------------------
    def find_paths(self, rots, poss, ends, max_actions: int = 5000):
        """Action lists that take every agent to its goal ([] where the follower fails); also
        returns the final positions."""
        rot = np.asarray(rots, np.float64).reshape(self.n, 4).copy()
        pos = np.asarray(poss, np.float64).reshape(self.n, 3).copy()
        end = np.asarray(ends, np.float64).reshape(self.n, 3)
        self.reset()
        running = np.ones(self.n, bool)
        while running.any():
            ix = np.nonzero(running)[0]
            prims = self._decide(rot[ix], pos[ix], end[ix])
            fwd = []
            for i, prim in zip(ix, prims):
                if not prim:
                    self._actions[i].append(GreedyFollowerCodes.ERROR)
                else:
                    for a in prim:
                        if a == GreedyFollowerCodes.RIGHT:
                            rot[i] = self._turn(rot[i], -1.0)
                        elif a == GreedyFollowerCodes.LEFT:
                            rot[i] = self._turn(rot[i], +1.0)
                        elif a == GreedyFollowerCodes.FORWARD:
                            fwd.append(i)
                        self._actions[i].append(a)
                if (self._actions[i][-1] in (GreedyFollowerCodes.STOP, GreedyFollowerCodes.ERROR)
                        or len(self._actions[i]) >= max_actions):
                    running[i] = False
            if fwd:  # every primitive ends with one FORWARD: one batched try_step for all of them
                fi = np.asarray(fwd)
                tgt = self._forward_target(rot[fi], pos[fi])
                pos[fi] = np.asarray(self.pathfinder.try_steps(pos[fi].astype(np.float32), tgt.astype(np.float32),
                                                               self.allow_sliding), np.float64)
        paths = []
        for i in range(self.n):
            a = self._actions[i]
            bad = a[-1] == GreedyFollowerCodes.ERROR or len(a) >= max_actions
            paths.append([] if bad else list(a))
        return paths, pos


class GreedyGeodesicFollowerImpl:
    """SPB.cpp:287-312 constructor order: (pathfinder, move_forward, turn_left, turn_right,
    goal_dist, forward_amount, turn_amount, fix_thrashing, thrashing_threshold).  The three
    MoveFn arguments are accepted for signature compatibility; the default kinematics are
    used, which is what enables batched evaluation.  A batch of one agent."""

    def __init__(self, pathfinder, move_forward=None, turn_left=None, turn_right=None,
                 goal_dist: float = 0.1875, forward_amount: float = 0.25,
                 turn_amount: float = math.radians(10.0), fix_thrashing: bool = True,
                 thrashing_threshold: int = 16):
        self.pathfinder = pathfinder
        self._b = GreedyGeodesicFollowerBatchImpl(pathfinder, 1, goal_dist, forward_amount, turn_amount,
                                                  fix_thrashing, thrashing_threshold)

    def reset(self):
        self._b.reset()

    def next_action_along(self, current_rot, current_pos, end):
        return self._b.next_actions_along(np.asarray(current_rot)[None], np.asarray(current_pos)[None],
                                          np.asarray(end)[None])[0]

    def find_path(self, current_rot, current_pos, end, max_actions: int = 5000):
        return self._b.find_paths(np.asarray(current_rot)[None], np.asarray(current_pos)[None],
                                  np.asarray(end)[None], max_actions)[0][0]


class GreedyGeodesicFollower:
    """habitat_sim.nav.GreedyGeodesicFollower twin that owns its (default) kinematics instead
    of an Agent: state is (rotation quaternion xyzw, position)."""

    def __init__(self, pathfinder, goal_radius=None, *, forward_amount: float = 0.25,
                 turn_degrees: float = 10.0, stop_key=None, forward_key="move_forward",
                 left_key="turn_left", right_key="turn_right", fix_thrashing: bool = True,
                 thrashing_threshold: int = 16):
        self.pathfinder = pathfinder
        self.goal_radius = 0.75 * forward_amount if goal_radius is None else goal_radius
        self.action_mapping = {GreedyFollowerCodes.STOP: stop_key,
                               GreedyFollowerCodes.FORWARD: forward_key,
                               GreedyFollowerCodes.LEFT: left_key,
                               GreedyFollowerCodes.RIGHT: right_key}
        self.impl = GreedyGeodesicFollowerImpl(pathfinder, None, None, None, self.goal_radius,
                                               forward_amount, math.radians(turn_degrees),
                                               fix_thrashing, thrashing_threshold)
        self.last_goal = None

    def reset(self):
        self.impl.reset()
        self.last_goal = None

    def next_action_along(self, rotation, position, goal_pos):
        if self.last_goal is None or not np.allclose(goal_pos, self.last_goal):
            self.reset()
            self.last_goal = np.asarray(goal_pos)
        act = self.impl.next_action_along(rotation, position, goal_pos)
        if act == GreedyFollowerCodes.ERROR:
            raise RuntimeError("GreedyFollowerError")
        return self.action_mapping[act]

    def find_path(self, rotation, position, goal_pos):
        self.reset()
        path = self.impl.find_path(rotation, position, goal_pos)
        if not path:
            raise RuntimeError("GreedyFollowerError")
        return [self.action_mapping[a] for a in path]


class GreedyGeodesicFollowerBatch:
    """N followers with the names of habitat_sim.nav.GreedyGeodesicFollower, one decision of all
    agents per call (habitat-lab's vectorised environments)."""

    def __init__(self, pathfinder, num_agents: int, goal_radius=None, *, forward_amount: float = 0.25,
                 turn_degrees: float = 10.0, stop_key=None, forward_key="move_forward",
                 left_key="turn_left", right_key="turn_right", fix_thrashing: bool = True,
                 thrashing_threshold: int = 16, allow_sliding: bool = True):
        self.goal_radius = 0.75 * forward_amount if goal_radius is None else goal_radius
        self.action_mapping = {GreedyFollowerCodes.STOP: stop_key,
                               GreedyFollowerCodes.FORWARD: forward_key,
                               GreedyFollowerCodes.LEFT: left_key,
                               GreedyFollowerCodes.RIGHT: right_key,
                               GreedyFollowerCodes.ERROR: None}
        self.impl = GreedyGeodesicFollowerBatchImpl(pathfinder, num_agents, self.goal_radius, forward_amount,
                                                    math.radians(turn_degrees), fix_thrashing,
                                                    thrashing_threshold, allow_sliding)
        self.last_goals = None

    def reset(self):
        self.impl.reset()
        self.last_goals = None

    def next_actions_along(self, rotations, positions, goal_positions):
        goals = np.asarray(goal_positions, np.float64).reshape(self.impl.n, 3)
        if self.last_goals is None:
            self.impl.reset()
        else:
            changed = [i for i in range(self.impl.n) if not np.allclose(goals[i], self.last_goals[i])]
            if changed:
                self.impl.reset(changed)
        self.last_goals = goals.copy()
        codes = self.impl.next_actions_along(rotations, positions, goals)
        return codes, [self.action_mapping[c] for c in codes]

    def find_paths(self, rotations, positions, goal_positions):
        paths, final = self.impl.find_paths(rotations, positions, goal_positions)
        return [[self.action_mapping[a] for a in p] for p in paths], final
